set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -k "tcgen05 or full_size" > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
grep -E "passed|failed|PASS|FAIL|Error|error" gpurun_out/pytest_tc.log | tail -40
