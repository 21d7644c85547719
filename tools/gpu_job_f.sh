python tools/quickbench.py --configs c4_coif5 --reps 3 --fuse 1 --colmin 3 > gpurun_out/ab_a.jsonl 2>&1
VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_b.so python tools/quickbench.py --configs c4_coif5 --reps 3 --fuse 1 --colmin 3 > gpurun_out/ab_b.jsonl 2>&1
cat gpurun_out/ab_a.jsonl gpurun_out/ab_b.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_column|k_fused" -c 20 --csv --log-file gpurun_out/kern_r1s.csv python tools/quickbench.py --configs c4_coif5 --reps 1 --fuse 1 --colmin 3 > /dev/null 2>&1
