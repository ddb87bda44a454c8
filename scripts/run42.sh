cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resident_order.py tests/test_gpu_pairs.py tests/test_gpu_domain.py tests/test_gpu_logic.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -20
python scripts/midsize_probe.py 2>&1 | tail -10
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'], d['block_list_policy'])"
