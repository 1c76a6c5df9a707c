set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --cpu-pairs 0 --matcher-layers 9 > gpurun_out/bench_lg9c.log 2>&1; tail -1 gpurun_out/bench_lg9c.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'matched', d['matched_fraction']); print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if k.startswith('lg_')})"
