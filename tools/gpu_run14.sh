set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "batch_path" 2>&1 | tail -30 > gpurun_out/pytest_lg2.log; tail -30 gpurun_out/pytest_lg2.log
timeout 600 python bench.py --steps 10 --cpu-pairs 0 --matcher-layers 9 > gpurun_out/bench_lg9.log 2>&1; tail -1 gpurun_out/bench_lg9.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'matched', d['matched_fraction']); print(json.dumps(d['kernel_ms_per_step'], indent=0))"
