cd /root/repo
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], (d.get('strip_parity') or {}).get('status'), (d.get('strip_kept_block_lists') or {}).get('status'), d.get('strip_phase_ms_rank0'), d['block_list_policy']['since_upload'])
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-2500:])
PY
}
run r2x_weak3_8 8 --steps 100 --warmup 10
run r2x_weak3_8_drv 8 --steps 20 --warmup 5
run r2x_strong16Mcirc_8 8 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
run r2x_weak3_4 4 --steps 100 --warmup 10
run r2x_strong16Mcirc_4 4 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
run r2x_strong4Mroom_4 4 --steps 50 --warmup 10 --agents 4000000 --scaling strong --workload room_exit
run r2x_strong4Mroom_2 2 --steps 50 --warmup 10 --agents 4000000 --scaling strong --workload room_exit
