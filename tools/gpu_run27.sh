mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "lightglue_layers_match_oracle or stereo_rotate or k2_keypoint" 2>&1 | tail -40 > gpurun_out/sanitizer_racecheck.log; grep -c "Race reported\|hazard" gpurun_out/sanitizer_racecheck.log; tail -8 gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "end_to_end or batch_equals or lightglue_layers_in_the_batch" 2>&1 | tail -6
