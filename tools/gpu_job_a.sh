set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1j.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_r1j.log
tail -3 gpurun_out/gpu_tests_r1j.log
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8,c4_coif5 --reps 5 > gpurun_out/quick_r1j.jsonl 2>&1
cat gpurun_out/quick_r1j.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 80 --csv --log-file gpurun_out/kern_r1j.csv python tools/quickbench.py --configs c4_coif5 --reps 1 > gpurun_out/ncu_r1j.log 2>&1
