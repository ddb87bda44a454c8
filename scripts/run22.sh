cd /root/repo
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], (d.get('strip_parity') or {}).get('status'), d.get('strip_phase_ms_rank0'))
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
run r2q_weak3_8 8 --steps 100 --warmup 10
run r2q_strong16Mcirc_8 8 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
run r2q_strong16Mcirc_4 4 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
run r2q_weak3_4 4 --steps 100 --warmup 10
run r2q_strong4Mroom_4 4 --steps 50 --warmup 10 --agents 4000000 --scaling strong --workload room_exit
