O=gpurun_out; mkdir -p $O
python tools/prof_once.py --wavelet sym8 --batch 1024 --log2n 16 --levels 8 --warm 1 > $O/plain_c3.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_lean|k_column" -s 8 -c 8 -o $O/p5_c3 -f \
   python tools/prof_once.py --wavelet sym8 --batch 1024 --log2n 16 --levels 8 --warm 1 > $O/ncu_c3.log 2>&1
ncu -i $O/p5_c3.ncu-rep --page source --csv > $O/p5_c3_source.csv 2>/dev/null
ls -la $O/p5*
