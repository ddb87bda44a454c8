cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resident_order.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{)|passed|failed|^FAILED" | cut -c1-300 | head -20
b() { name=$1; lib=$2; shift 2; CROWD_B200_LIB=$lib python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()})
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
V=/root/repo/crowddynamics_b200/csrc/variants
for m in three_circle circular; do
for v in A B C D E; do b r2u_${m}_$v $V/lib_$v.so --model $m; done
done
