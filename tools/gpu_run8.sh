set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused|nms_r4" -s 6 -c 2 -o gpurun_out/prof_fused_nms_r1 python bench.py --steps 1 --warmup 3 --batch 8 --cpu-pairs 0 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
