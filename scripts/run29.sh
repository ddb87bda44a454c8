cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED" | cut -c1-300 | head -20
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], (d.get('strip_parity') or {}).get('status'), (d.get('strip_kept_block_lists') or {}), d.get('strip_phase_ms_rank0'), d['block_list_policy'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-2500:])
PY
}
run r2w_weak3_2 2 --steps 100 --warmup 10
run r2w_weak3_2_K1 2 --steps 100 --warmup 10 --rebuild-max 1
run r2w_weakcirc_2 2 --steps 100 --warmup 10 --model circular
run r2w_strong16Mcirc_2 2 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
