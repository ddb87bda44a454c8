cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resident_order.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{)|passed|failed|^FAILED" | cut -c1-400 | head -60
