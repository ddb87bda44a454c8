for n in 2 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/scale_three_circle_$n.json
  python -c "import json; d=json.load(open('gpurun_out/scale_three_circle_$n.json')); print('three_circle', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tests/run_strips_nccl.py 2>&1 | grep strips
