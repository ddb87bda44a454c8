cd /root/repo
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_logic.py tests/test_gpu_parity.py tests/test_gpu_domain.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -30
python bench.py --workload hallway > gpurun_out/r2y_hallway.json 2>gpurun_out/r2y_hallway.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_hallway.json').read().strip().splitlines()[-1]); print({k:round(v['us_per_update'],1) for k,v in d['hallway'].items()})
PY
