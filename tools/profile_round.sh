#!/bin/bash
# Produces the round's measurement artefacts under gpurun_out/ (copied into profiles/ by tools/profile_collect.py):
#   1. plain bench line (no profiler)          2. ncu launch list of the same command (gpu__time_duration per launch)
#   3. one `ncu --set full` capture of the dominant kernel of the default workload and of the coif5 long-signal kernels
# Usage (on the GPU box): bash tools/profile_round.sh r01 [bench|coif5]   -- one part per gpurun call: the two .ncu-rep
# files together exceed the 64 MiB that gpurun copies back
R=${1:-r01}
PART=${2:-bench}
O=gpurun_out
mkdir -p $O
if [ "$PART" = bench ]; then
python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err || { echo "bench failed"; tail -5 $O/bench_$R.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches_$R.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_fused_(analysis|synthesis)" --launch-skip 6 -c 2 \
    -o $O/prof_${R}_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_$R.log 2>&1
bash tools/gpu_results.sh $R > /dev/null 2>&1
else
python tools/prof_once.py --warm 0 > /dev/null 2>&1 && \
# (ten forward launches + the first four inverse ones; no source import: gpurun_out/ is capped at 64 MiB per call)
timeout 900 ncu --set full --clock-control none -k regex:"k_column|k_fused" -c 14 -o $O/prof_${R}_coif5 -f \
    python tools/prof_once.py --warm 0 > $O/ncu_full_coif5_$R.log 2>&1
fi
ls -la $O/*.ncu-rep
