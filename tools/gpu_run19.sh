set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_ -s 33 -c 11 -o gpurun_out/prof_lg_r3 python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 --matcher-layers 1 > gpurun_out/ncu_lg3.log 2>&1
ls -la gpurun_out/*.ncu-rep
