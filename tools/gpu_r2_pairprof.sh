# ncu full capture of the eight pair launches of config #4 at LOG2N (default 2^25: one rank's span of an 8-GPU job)
O=gpurun_out; mkdir -p $O
L=${1:-25}
python tools/prof_once.py --warm 0 --log2n $L > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lat2" -c 8 -o $O/prof_r02_pair$L -f \
    python tools/prof_once.py --warm 0 --log2n $L > $O/ncu_pair.log 2>&1
python tools/ncu_summary.py $O/prof_r02_pair$L.ncu-rep
