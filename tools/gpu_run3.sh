set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -3 gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --cpu-pairs 4 > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
