#!/usr/bin/env bash
# The reproducible GPU recipes of this repo, run ON a B200 box:   gpurun --timeout 1500 -- 'bash tools/gpu.sh <recipe> [args]'
# Everything a recipe writes goes to gpurun_out/ (merged back by gpurun); summaries worth keeping are copied to profiles/.
#
#   tests                      python -m pytest tests -m gpu
#   bench [bench.py args]      the default bench line (config 3, both precisions) -> gpurun_out/bench.json
#   reference                  the CPU reference arm -> gpurun_out/bench_reference.json
#   launches <precision>       ncu launch list (gpu__time_duration.sum) of two steps -> gpurun_out/launches_<precision>.csv
#   ncu <precision> [batch]    ncu --set full of ONE step (batch 16 by default) summarised by tools/ncu_summary.py
#                              -> gpurun_out/ncu_step_<precision>.{json,md}
#   ncu_kernel <regex> [precision] [skip]   ncu --set full with source of ONE launch of a kernel -> gpurun_out/ncu_<regex>_{source,raw}.csv, _details.txt
#   sanitizer                  compute-sanitizer memcheck + racecheck over a small batch call
#   stream | config5           bench.py --config 4 / 5 on the GPUs of this box (torchrun when --gpus > 1: use run_n)
#   run_n <N> [bench.py args]  torchrun launch exactly like the driver's (N ranks on one box)
set -u
mkdir -p gpurun_out
recipe=${1:-tests}; shift || true
case "$recipe" in
  tests)     timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log ;;
  bench)     timeout 1400 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json ;;
  reference) timeout 1400 python bench.py --impl reference --steps 5 --warmup 1 "$@" | tee gpurun_out/bench_reference.json ;;
  launches)  p=${1:-fp32_faithful}
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$p.csv \
               python bench.py --steps 2 --warmup 3 --batch 16 --cpu-pairs 0 --acc-pairs 0 --precision $p > gpurun_out/ncu_launches.log 2>&1
             grep -c . gpurun_out/launches_$p.csv ;;
  ncu)       p=${1:-fp32_faithful}; b=${2:-16}
             # one step = the launches between two finalize_kernel launches; take the 4th step (after the warm-up) from the launch list
             [ -f gpurun_out/launches_$p.csv ] || bash tools/gpu.sh launches $p
             read skip count < <(python - "$p" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{sys.argv[1]}.csv")) if len(r) > 5]
k = rows[0].index("Kernel Name")
fin = [i for i, r in enumerate(rows[1:]) if "finalize_kernel" in r[k]]
print(fin[2] + 1, fin[3] - fin[2])
PY
)
             echo "ncu --set full: skip $skip launches, capture $count"
             timeout 1400 ncu --set full --clock-control none --import-source on --launch-skip $skip --launch-count $count -o /tmp/prof_step_$p \
               python bench.py --steps 1 --warmup 3 --batch $b --cpu-pairs 0 --acc-pairs 0 --precision $p > gpurun_out/ncu_step.log 2>&1
             python tools/ncu_summary.py /tmp/prof_step_$p.ncu-rep gpurun_out/ncu_step_$p.json > gpurun_out/ncu_step_$p.md; tail -3 gpurun_out/ncu_step_$p.md ;;
  ncu_kernel) k=$1; p=${2:-bf16_fast}; skip=${3:-8}
             timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -o /tmp/k_$k \
               python bench.py --steps 1 --warmup 3 --batch 16 --cpu-pairs 0 --acc-pairs 0 --sift-pairs 0 --precision $p > gpurun_out/ncu_kernel.log 2>&1
             ncu -i /tmp/k_$k.ncu-rep --page source --csv > gpurun_out/ncu_${k}_source.csv 2>/dev/null
             ncu -i /tmp/k_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${k}_raw.csv 2>/dev/null
             ncu -i /tmp/k_$k.ncu-rep --page details > gpurun_out/ncu_${k}_details.txt 2>/dev/null
             ls -la gpurun_out/ncu_${k}_* ;;
  sanitizer) for tool in memcheck racecheck; do
               timeout 700 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
                 -k "batch_equals_single or two_stream or lazy_rasters or k2_keypoint or list_overflow or x3_layers" \
                 > gpurun_out/sanitizer_$tool.log 2>&1; tail -4 gpurun_out/sanitizer_$tool.log; done ;;
  stream)    timeout 1400 python bench.py --config 4 --steps 200 "$@" | tee gpurun_out/bench_config4.json ;;
  config5)   timeout 1400 python bench.py --config 5 --steps 4 "$@" | tee gpurun_out/bench_config5.json ;;
  run_n)     n=$1; shift
             timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
               bench.py --gpus $n "$@" 2> gpurun_out/run_n$n.err | tee gpurun_out/run_n$n.json; tail -2 gpurun_out/run_n$n.err ;;
  *) echo "unknown recipe $recipe"; exit 2 ;;
esac
