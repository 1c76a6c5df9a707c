set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -k "k2 or end_to_end or full_size" > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k2.log
tail -3 gpurun_out/pytest_k2.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac']); print({k: round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})"
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
if [ "$NG" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --batch 8 > gpurun_out/bench_n$NG.log 2>&1; tail -1 gpurun_out/bench_n$NG.log | cut -c1-600
fi
