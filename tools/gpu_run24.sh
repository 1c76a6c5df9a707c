mkdir -p gpurun_out
GNB_CONV_DEBUG=1 timeout 600 python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 2>&1 | grep F1_DBG | tail -4
