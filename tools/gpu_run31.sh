mkdir -p gpurun_out
timeout 900 python bench.py --batch 64 --keypoints 2048 --ransac-iters 2000 --steps 10 --cpu-pairs 0 > gpurun_out/bench_config5_n1.log 2>&1; tail -1 gpurun_out/bench_config5_n1.log | cut -c1-900
