set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 105 --launch-count 70 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --launch-skip 105 --launch-count 35 -o /tmp/prof_step_final python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_step.log 2>&1
python tools/ncu_summary.py /tmp/prof_step_final.ncu-rep gpurun_out/ncu_full_step_batch16.json > gpurun_out/ncu_full_step_batch16.md
tail -3 gpurun_out/ncu_full_step_batch16.md
