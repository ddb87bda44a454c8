cd /root/repo
timeout 900 python -m pytest tests/test_gpu_strips.py -q -m gpu 2>&1 | tail -2
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r3d_weak3_2.json 2> gpurun_out/r3d_weak3_2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d_weak3_2.json').read().strip().splitlines()[-1]); print('%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['strip_parity']['status'], d['strip_kept_block_lists']['status'], d['strip_phase_ms_rank0'])
PY
done
