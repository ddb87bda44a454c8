cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_field.py -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --workload room_exit --agents 4000000 --field-step 0.1 --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2k_room4M.json 2> gpurun_out/r2k_room4M.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_room4M.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['field_build'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2k_room4M.err').read()[-2000:])
PY
