timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_column|k_fused" -c 40 --csv --log-file gpurun_out/kern_r1r.csv python tools/quickbench.py --configs c4_haar,c4_sym8 --reps 1 --fuse 1 --colmin 3 --poly 2 > gpurun_out/r1r.log 2>&1
tail -2 gpurun_out/r1r.log
