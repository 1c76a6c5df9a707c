GNB_ATTN_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --batch 16 --cpu-pairs 0 --matcher-layers 1 2>&1 | grep FC1_DBG | tail -3
