set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stereo" 2>&1 | tail -30 > gpurun_out/pytest_stereo.log; tail -30 gpurun_out/pytest_stereo.log
