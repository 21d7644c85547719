python tools/quickbench.py --configs c4_coif5,c3_sym8 --reps 3 --fuse 1 --colmin 3 > gpurun_out/ab_a.jsonl 2>&1
VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_b.so python tools/quickbench.py --configs c4_coif5,c3_sym8 --reps 3 --fuse 1 --colmin 3 > gpurun_out/ab_b.jsonl 2>&1
cat gpurun_out/ab_a.jsonl gpurun_out/ab_b.jsonl
