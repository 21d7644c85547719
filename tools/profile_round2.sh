#!/bin/bash
# Round-2 measurement artefacts under gpurun_out/ (tools/profile_collect2.py turns them into profiles/r02_*):
#   1. the plain bench line (no profiler) and the reference-arm line
#   2. ncu launch list of the bench command (gpu__time_duration per launch)
#   3. one `ncu --set full` capture (source imported) of the dominant kernels of the default workload
#   4. DRAM bytes + duration of every engine kernel of ONE forward + inverse (+ denoise) of each BASELINE config
# bash tools/profile_round2.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err || { echo "bench failed"; tail -5 $O/bench_$R.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2> /dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_launches_$R.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_lean_(analysis|synthesis)" --launch-skip 6 -c 2 \
    -o $O/prof_${R}_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_full_$R.log 2>&1
ncu -i $O/prof_${R}_bench.ncu-rep --page source --csv > $O/prof_${R}_bench_source.csv 2>/dev/null
# config #4's kernels at a quarter of its length: the two coif5 tile levels and the lattice column pairs (one launch of each kind per direction)
python tools/prof_once.py --warm 0 --log2n 26 > /dev/null 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"^k_|::k_" -c 12 -o $O/prof_${R}_coif5 -f \
    python tools/prof_once.py --warm 0 --log2n 26 > $O/ncu_full_coif5_$R.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
cap() {  # name, prof_once args...
  name=$1; shift
  python tools/prof_once.py --warm 0 "$@" > /dev/null 2>&1 &&
  timeout 900 ncu --metrics $M --clock-control none -k regex:"^k_|::k_" --csv --log-file $O/traffic_${R}_$name.csv \
      python tools/prof_once.py --warm 0 "$@" > $O/traffic_${R}_$name.log 2>&1
}
cap batch4096x4096_db4_J4 --wavelet db4 --batch 4096 --log2n 12 --levels 4
cap batch4096x4096_haar_J4 --wavelet haar --batch 4096 --log2n 12 --levels 4
cap batch1024x65536_sym8_J8 --wavelet sym8 --batch 1024 --log2n 16 --levels 8
cap batch256x1M_db8_J6 --wavelet db8 --batch 256 --log2n 20 --levels 6
cap c5_ZERO_PADDING --wavelet db8 --batch 256 --log2n 20 --levels 6 --mode 1 --denoise 1
cap c5_SYMMETRIC --wavelet db8 --batch 256 --log2n 20 --levels 6 --mode 2 --denoise 1
cap single2p28_coif5_J10 --wavelet coif5 --batch 1 --log2n 28 --levels 10
ls -la $O/*.ncu-rep $O/traffic_${R}_*.csv
