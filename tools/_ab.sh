GNB_SIDE_STREAM=1 timeout 600 python -m pytest tests/test_gpu_precise.py -q -m gpu -x -k "end_to_end or composes" 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 500 python bench.py --steps 8 --warmup 3 --cpu-pairs 0 --acc-pairs 8 --sift-pairs 0 --precision fp32_faithful > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; tail -2 gpurun_out/ab_$tag.err; }
run side GNB_SIDE_STREAM=1
run noside GNB_SIDE_STREAM=0
run side2 GNB_SIDE_STREAM=1
