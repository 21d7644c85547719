set -x
python tools/prof_once.py --warm 0 > gpurun_out/prof_once.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_column|k_fused" -c 14 -o gpurun_out/prof_r1k_c4 -f python tools/prof_once.py --warm 0 > gpurun_out/ncu_r1k.log 2>&1
ls -la gpurun_out/*.ncu-rep
