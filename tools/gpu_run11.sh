set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cpu-pairs 2 > gpurun_out/bench_r11.log 2>&1; tail -1 gpurun_out/bench_r11.log | cut -c1-3000
