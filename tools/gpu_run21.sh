set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/scale_n8.log 2>&1; tail -1 gpurun_out/scale_n8.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/scale_ref_n8.log 2>&1; tail -2 gpurun_out/scale_ref_n8.log | cut -c1-300
