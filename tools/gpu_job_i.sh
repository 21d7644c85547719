python -m pytest tests -m gpu -x -q -k "denoise or swt or thresh or median or column" 2>&1 | tail -3
for m in 1 2 0; do python tools/quickbench.py --configs c5_db8 --reps 3 --mode $m --denoise 1; done > gpurun_out/quick_r1q.jsonl 2>&1
cat gpurun_out/quick_r1q.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/kern_r1q.csv python tools/quickbench.py --configs c5_db8 --reps 1 --mode 2 --denoise 1 > /dev/null 2>&1
