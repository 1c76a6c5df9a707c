set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --cpu-pairs 4 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-200
timeout 600 python bench.py --steps 20 --cpu-pairs 0 --matcher-layers 9 > gpurun_out/bench_layers9.log 2>&1; tail -1 gpurun_out/bench_layers9.log | cut -c1-200
timeout 600 python tools/bench_next_rows.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; tail -3 gpurun_out/next_rows.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 105 --launch-count 70 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --launch-skip 105 --launch-count 35 -o /tmp/prof_step_final python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 > gpurun_out/ncu_step.log 2>&1
python tools/ncu_summary.py /tmp/prof_step_final.ncu-rep gpurun_out/ncu_full_step_batch16.json > gpurun_out/ncu_full_step_batch16.md
timeout 900 ncu --set full --clock-control none -k regex:lg_ -s 33 -c 11 -o /tmp/prof_lg_final python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 --matcher-layers 1 > gpurun_out/ncu_lg_final.log 2>&1
python tools/ncu_summary.py /tmp/prof_lg_final.ncu-rep gpurun_out/ncu_lightglue_layer_batch16.json > gpurun_out/ncu_lightglue_layer_batch16.md
du -sh gpurun_out
