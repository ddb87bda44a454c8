cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resident_order.py tests/test_gpu_domain.py tests/test_gpu_logic.py tests/test_gpu_pairs.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -20
