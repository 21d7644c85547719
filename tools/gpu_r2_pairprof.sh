# ncu full capture of the lattice column kernels of config #4 at 2^26 (one analysis + one synthesis launch of each kind)
O=gpurun_out; mkdir -p $O
python tools/prof_once.py --warm 0 --log2n 26 > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lat" --launch-skip 3 -c 2 -o $O/prof_r02_pair -f \
    python tools/prof_once.py --warm 0 --log2n 26 > $O/ncu_pair.log 2>&1
python tools/ncu_summary.py $O/prof_r02_pair.ncu-rep
