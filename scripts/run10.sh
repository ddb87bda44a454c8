cd /root/repo
for m in three_circle circular; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --model $m > gpurun_out/r2i_${m}_2.json 2> gpurun_out/r2i_${m}_2.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i_${m}_2.json').read().strip().splitlines()[-1]); print('$m x2', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2i_${m}_2.err').read()[-1500:])
PY
done
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2i_three_circle_1.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r2i_three_circle_1.json').read().strip().splitlines()[-1]); print('x1', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'])"
