# round-2 first box call: JDK probe, GPU suite, queued A/B variants
O=gpurun_out; mkdir -p $O
{ echo "== jdk probe"; command -v java javac jshell; ls /usr/lib/jvm /opt 2>&1; find / -name "javac*" -not -path "/proc/*" 2>/dev/null | head; nproc; nvidia-smi -L; } > $O/jdk_probe.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in b c d; do
  echo "== variant $v"
  VW_LIB_PATH=/root/repo/vectorwave_b200/libvwmodwt_$v.so timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
  bash tools/ab_libs.sh /root/repo/vectorwave_b200/libvwmodwt_$v.so c2_haar,c2_db4
done
