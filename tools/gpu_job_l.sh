nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 20 > gpurun_out/smi_c4.csv &
SMI=$!
python tools/quickbench.py --configs c4_coif5 --reps 20 --fuse 1 --colmin 3 > gpurun_out/c4_clk.jsonl 2>&1
kill $SMI
sort gpurun_out/smi_c4.csv | uniq -c | sort -rn | head -12
cat gpurun_out/c4_clk.jsonl | cut -c1-300
