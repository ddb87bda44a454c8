cd /root/repo
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_driver.json 2> gpurun_out/r2z_driver.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_driver.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['roofline_fp64'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
PY
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 | tail -c 600
