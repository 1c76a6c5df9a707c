mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "candidate" 2>&1 | tail -15
