set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --launch-skip 105 --launch-count 35 -o gpurun_out/prof_step_r1 python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 105 --launch-count 70 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_launches.log 2>&1
timeout 900 python bench.py --cpu-pairs 4 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-400
ls -la gpurun_out/*.ncu-rep
