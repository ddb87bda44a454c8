cd /root/repo
b() { name=$1; lib=$2; shift 2; CROWD_B200_LIB=$lib python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()})
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
V=/root/repo/crowddynamics_b200/csrc
b r3a_three_base $V/libcrowd_b200.so
for v in F G H I; do b r3a_three_$v $V/variants/lib_$v.so; done
b r3a_circ_base $V/libcrowd_b200.so --model circular
for v in F G H I; do b r3a_circ_$v $V/variants/lib_$v.so --model circular; done
