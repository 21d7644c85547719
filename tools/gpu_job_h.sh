python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1p.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_r1p.log
tail -15 gpurun_out/gpu_tests_r1p.log
python tools/quickbench.py --configs c5_db8 --reps 3 --mode 2 > gpurun_out/quick_r1p.jsonl 2>&1
python tools/quickbench.py --configs c5_db8 --reps 3 --mode 1 >> gpurun_out/quick_r1p.jsonl 2>&1
python tools/quickbench.py --configs c2_db4,c2_haar,c3_sym8 --reps 10 >> gpurun_out/quick_r1p.jsonl 2>&1
cat gpurun_out/quick_r1p.jsonl
