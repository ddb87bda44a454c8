cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | head
echo "== memcheck strips (circular world-2 cases) + pairs"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_strips.py tests/test_gpu_pairs.py -q -m gpu -k "not nccl and not dts1 and not 3-" 2>&1 | grep -v "^E  " | tail -25 > gpurun_out/memcheck_r2.txt; tail -12 gpurun_out/memcheck_r2.txt
