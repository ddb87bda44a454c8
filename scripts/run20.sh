cd /root/repo
for f in collective domain edge_cases field fluctuation fullsize logic parity transfers; do
  echo "== $f + strips"; timeout 900 python -m pytest tests/test_gpu_$f.py tests/test_gpu_strips.py -q -m gpu 2>&1 | grep -E "passed|failed|^FAILED" | head -12
done
