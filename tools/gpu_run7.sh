set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
for N in 1 2 4 8; do
  if [ "$N" -le "$NG" ]; then
    if [ "$N" -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n1.log 2>&1
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
    fi
    tail -1 gpurun_out/scale_n$N.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'matched', d['matched_fraction'], 'rmse_gt', d['pose_rmse_px_vs_ground_truth'], 'clocks', d['clocks'])"
  fi
done
