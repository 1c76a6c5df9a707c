set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lightglue" 2>&1 | tail -40 > gpurun_out/pytest_lg.log; tail -40 gpurun_out/pytest_lg.log
