cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resident_order.py -q -m gpu -x 2>&1 | grep -v "^E   " | tail -25
b() { name=$1; shift; python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()}, d['block_list_policy']['since_upload'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
b r2r_three_K1 --rebuild-max 1
b r2r_three_K16 --rebuild-max 16
b r2r_three_K16_s05 --rebuild-max 16 --skin 0.05
b r2r_three_K32_s15 --rebuild-max 32 --skin 0.15
b r2r_three_K16_fine --rebuild-max 16 --refinement 2
b r2r_three_K1_fine --rebuild-max 1 --refinement 2
b r2r_circ_K1 --model circular --rebuild-max 1
b r2r_circ_K16 --model circular --rebuild-max 16
b r2r_circ_K16_s05 --model circular --rebuild-max 16 --skin 0.05
