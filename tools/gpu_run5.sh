set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/bench4.log 2>&1; tail -1 gpurun_out/bench4.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac']); print(json.dumps(d['kernel_ms_per_step'], indent=0))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -s 24 -c 2 -o gpurun_out/prof_conv_halo_r1 python bench.py --steps 1 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
