cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py -q -m gpu -k "kept and 2-circular and messages and adaptive" 2>&1 | grep -B6 "^E " | cut -c1-400 | head -40
