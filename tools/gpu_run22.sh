set -x
mkdir -p gpurun_out
for N in 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.log 2>&1; tail -1 gpurun_out/scale_n$N.log | cut -c1-200
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --cpu-pairs 0 > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | cut -c1-200
