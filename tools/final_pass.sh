#!/usr/bin/env bash
# One GPU call that refreshes what lives under profiles/: tests, the default bench line, launch lists + ncu step summaries
# of both precisions, the config 1 / 2 / 4 lines, sanitizer logs.  Usage: gpurun -- 'bash tools/final_pass.sh [part ...]'
# parts: tests bench ncu configs sanitizer (default: all)
set -u
mkdir -p gpurun_out
parts=${*:-tests bench ncu configs sanitizer}
for part in $parts; do
  case "$part" in
    tests)   bash tools/gpu.sh tests ;;
    bench)   bash tools/gpu.sh bench ;;
    ncu)     for p in fp32_faithful bf16_fast; do rm -f gpurun_out/launches_$p.csv; bash tools/gpu.sh ncu $p; done ;;
    configs) for c in 1 2; do timeout 600 python bench.py --config $c > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err; tail -1 gpurun_out/bench_config$c.err; done
             timeout 600 python bench.py --config 4 --steps 200 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.err
             timeout 600 python bench.py --config 4 --steps 200 --precision bf16_fast > gpurun_out/bench_config4_fast.json 2> gpurun_out/bench_config4_fast.err ;;
    sanitizer) bash tools/gpu.sh sanitizer ;;
  esac
done
ls gpurun_out | wc -l
