cd /root/repo
TAG=r2z bash profiles/capture_r2.sh > gpurun_out/capture_r2z.log 2>&1
python bench.py > gpurun_out/r2z_three_circle.json 2> gpurun_out/r2z_three_circle.err
python bench.py --model circular > gpurun_out/r2z_circular.json 2> gpurun_out/r2z_circular.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_three_circle_driver.json 2>/dev/null
python bench.py --model circular --agents 16000000 --steps 60 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2z_circular_16M.json 2>/dev/null
python bench.py --workload room_exit --agents 4000000 --steps 60 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2z_room_exit_4M.json 2>/dev/null
python bench.py --workload hallway > gpurun_out/r2z_hallway.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_*.json')):
    if 'dense' in f or 'driver.json'==f: continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.4g'%d['value'], '%.4f'%d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('traffic'), (d.get('roofline') or {}).get('phase_ms_per_step'))
        if 'hallway' in d: print({k:round(v['us_per_update'],1) for k,v in d['hallway'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
