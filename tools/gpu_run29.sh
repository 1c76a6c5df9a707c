timeout 600 python bench.py --steps 10 --cpu-pairs 0 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if k.startswith('conv')})"
