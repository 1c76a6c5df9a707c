set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/bench3.log 2>&1; tail -1 gpurun_out/bench3.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac']); print(json.dumps(d['kernel_ms_per_step'], indent=0))"
