R=${1:-r01}
python tools/quickbench.py --configs c2s_db4 --reps 50 > gpurun_out/results_$R.jsonl 2>&1
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c4_coif5 --reps 10 >> gpurun_out/results_$R.jsonl 2>&1
for m in 0 1 2; do python tools/quickbench.py --configs c5_db8 --reps 5 --mode $m --denoise 1; done >> gpurun_out/results_$R.jsonl 2>&1
cat gpurun_out/results_$R.jsonl | cut -c1-420
