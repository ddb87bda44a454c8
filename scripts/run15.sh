cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py -q -m gpu -x 2>&1 | tail -8
for x in peer nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --exchange $x > gpurun_out/r2m_three_2_$x.json 2> gpurun_out/r2m_three_2_$x.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2m_three_2_$x.json').read().strip().splitlines()[-1]); print('$x x2', d['value'], d['ms_per_step'], d['strip_exchange'], d['strip_parity']['status'], d['strip_phase_ms_rank0'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2m_three_2_$x.err').read()[-2500:])
PY
done
