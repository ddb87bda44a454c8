cd /root/repo
echo "== memcheck: resident-order, small crowds, strips with kept lists (reduced sizes via -k)"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_small.py tests/test_gpu_resident_order.py tests/test_gpu_strips.py -q -m gpu -k "(small and (50 or 7- or 256)) or (kept_order_equals and 8-steps and fixed) or stale_lattice or (kept_block_lists and 2- and fixed) or deferred" 2>&1 | grep -v "^E  " | tail -12 > gpurun_out/memcheck_r2z.txt; tail -8 gpurun_out/memcheck_r2z.txt
echo "== racecheck: small-crowd kernel (shared memory), finish kernel reductions"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_small.py tests/test_gpu_resident_order.py -q -m gpu -k "(small_kernel_equals and (50 or 256)) or (kept_order_equals and 8-steps and fixed and circular)" 2>&1 | grep -v "^E  " | tail -12 > gpurun_out/racecheck_r2z.txt; tail -8 gpurun_out/racecheck_r2z.txt
