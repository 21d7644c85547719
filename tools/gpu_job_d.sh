python -m pytest tests -m gpu -x -q -k "column or deep" 2>&1 | tail -2
for opts in "--fuse 1 --colmin 3" "--fuse 1 --colmin 4" "--fuse 1 --colmin 6" "--fuse 2 --colmin 3" "--fuse 0 --colmin 0"; do python tools/quickbench.py --configs c4_coif5 --reps 3 $opts; done > gpurun_out/sweep_r1m.jsonl 2>&1
cat gpurun_out/sweep_r1m.jsonl
