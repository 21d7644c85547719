timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_column" --launch-skip 3 -c 2 -o gpurun_out/prof_r1n_col -f python tools/prof_once.py --warm 0 > gpurun_out/ncu_r1n.log 2>&1
ls -la gpurun_out/prof_r1n_col.ncu-rep
