set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --cpu-pairs 4 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-120
timeout 600 python bench.py --steps 20 --cpu-pairs 0 --matcher-layers 9 > gpurun_out/bench_layers9.log 2>&1; tail -1 gpurun_out/bench_layers9.log | cut -c1-200
timeout 600 python tools/bench_next_rows.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; tail -2 gpurun_out/next_rows.err
