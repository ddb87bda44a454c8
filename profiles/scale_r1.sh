# scaling run: bench.py at N = 2, 4, 8 (three_circle, 1 M agents per GPU) and the north-star config (16 M circular on 8 GPUs)
mkdir -p gpurun_out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/scale_three_circle_$n.json
  python -c "import json; d=json.load(open('gpurun_out/scale_three_circle_$n.json')); print('three_circle', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --model circular --agents 2000000 --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/scale_circular16M_8.json
python -c "import json; d=json.load(open('gpurun_out/scale_circular16M_8.json')); print('circular 16M', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tests/run_strips_nccl.py 2>&1 | grep strips
