cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py tests/test_gpu_fluctuation.py -q -m gpu -x 2>&1 | tail -8
for m in three_circle circular; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --model $m > gpurun_out/r2n_${m}_2.json 2> gpurun_out/r2n_${m}_2.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n_${m}_2.json').read().strip().splitlines()[-1]); print('$m x2', d['value'], d['ms_per_step'], d['strip_parity']['status'], d['strip_phase_ms_rank0'], d['roofline']['phase_ms_per_step'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2n_${m}_2.err').read()[-2500:])
PY
done
