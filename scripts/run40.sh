cd /root/repo
b() { name=$1; shift; python bench.py --no-cpu-baseline --no-fp64-peak --e2e-steps 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.2f us/step'%(1e3*d['ms_per_step']), d['gpu_launches'], d['block_list_policy']['since_upload'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-800:])
PY
}
for n in 1000 4000 16000; do
for m in circular three_circle; do
b r3b_${m}_${n}_graph --model $m --agents $n --steps 2000 --warmup 20
b r3b_${m}_${n}_chain --model $m --agents $n --steps 2000 --warmup 20 --rebuild-min-agents 0
done; done
