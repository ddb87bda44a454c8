cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_field.py tests/test_gpu_transfers.py -q -m gpu -x 2>&1 | tail -25
