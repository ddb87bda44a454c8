cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py tests/test_gpu_domain.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -30
