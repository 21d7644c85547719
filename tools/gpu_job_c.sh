set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1l.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_r1l.log
tail -3 gpurun_out/gpu_tests_r1l.log
python tools/quickbench.py --configs c3_sym8,c5_db8,c4_coif5 --reps 5 > gpurun_out/quick_r1l.jsonl 2>&1
cat gpurun_out/quick_r1l.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_column|k_fused" -c 14 --csv --log-file gpurun_out/kern_r1l.csv python tools/prof_once.py --warm 0 > gpurun_out/ncu_r1l.log 2>&1
