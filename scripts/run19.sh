cd /root/repo
echo "== alone"; timeout 600 python -m pytest tests/test_gpu_strips.py -q -m gpu 2>&1 | tail -5
echo "== after others"; timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_random_sweep.py tests/test_gpu_strips.py -q -m gpu 2>&1 | grep -v "^E  " | tail -15
echo "== initcheck"; timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_gpu_strips.py -q -m gpu -k "test_strips_reproduce_single_device and messages and dts0 and circular and not three" 2>&1 | grep -v "^E  " | tail -80 > gpurun_out/initcheck_strips.txt; tail -60 gpurun_out/initcheck_strips.txt
