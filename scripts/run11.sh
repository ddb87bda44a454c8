cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2j_three.json 2> gpurun_out/r2j_three.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2j_three.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']); [print(k, v['value'], v['h2d_bytes_per_step'], v['d2h_bytes_per_step']) for k,v in d['e2e_variants'].items()]
PY
tail -5 gpurun_out/r2j_three.err
