mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "k1_dense" 2>&1 | tail -6
