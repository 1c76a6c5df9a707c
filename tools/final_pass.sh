#!/usr/bin/env bash
# One GPU call that refreshes everything under profiles/ that depends on the kernels: tests, default bench line,
# launch lists + ncu step summaries of both precisions, configs 1 / 2 / 4 lines.  Usage: gpurun -- 'bash tools/final_pass.sh'
set -u
mkdir -p gpurun_out
bash tools/gpu.sh tests
bash tools/gpu.sh bench
for p in fp32_faithful bf16_fast; do rm -f gpurun_out/launches_$p.csv; bash tools/gpu.sh ncu $p; done
for c in 1 2; do timeout 600 python bench.py --config $c > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err; tail -1 gpurun_out/bench_config$c.err; done
timeout 600 python bench.py --config 4 --steps 200 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.err
ls -la gpurun_out | tail -20
