set -x
mkdir -p gpurun_out
for B in 8 16 32; do
timeout 600 python bench.py --steps 20 --warmup 3 --batch $B --cpu-pairs 0 > gpurun_out/bench_b$B.log 2>&1; tail -1 gpurun_out/bench_b$B.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('batch', d['config']['pairs_per_step_per_gpu'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), 'clk', d['clocks'])"
done
