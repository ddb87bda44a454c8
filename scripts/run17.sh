cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_strips.py -q -m gpu -x 2>&1 | tail -4
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  if [ $np = 1 ]; then python bench.py --gpus 1 "$@" --no-cpu-baseline --e2e-steps 1 > gpurun_out/$name.json 2> gpurun_out/$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], (d.get('strip_parity') or {}).get('status'), d.get('strip_phase_ms_rank0'))
except Exception as e:
    print('$name ERR', e); print(open('gpurun_out/$name.err').read()[-1500:])
PY
}
run r2o_weak3_2 2 --steps 100 --warmup 10
run r2o_strong16Mcirc_2 2 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong
run r2o_strong4Mroom_2 2 --steps 50 --warmup 10 --agents 4000000 --scaling strong --workload room_exit
run r2o_strong16Mcirc_1 1 --steps 30 --warmup 5 --model circular --agents 16000000
