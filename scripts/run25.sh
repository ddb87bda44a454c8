cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{)|passed|failed|^FAILED" | cut -c1-300 | head -30
b() { name=$1; shift; python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()}, d['block_list_policy']['since_upload'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
b r2s_three
b r2s_three_K1 --rebuild-max 1
b r2s_circ --model circular
b r2s_circ16M --model circular --agents 16000000 --steps 50
b r2s_room4M --workload room_exit --agents 4000000 --steps 50
