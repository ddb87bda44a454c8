cd /root/repo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3c_scale_1.json 2> gpurun_out/r3c_scale_1.err
for np in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 > gpurun_out/r3c_scale_$np.json 2> gpurun_out/r3c_scale_$np.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 10 --model circular --agents 16000000 --scaling strong > gpurun_out/r3c_strong16M_8.json 2> gpurun_out/r3c_strong16M_8.err
python - <<'PY'
import json
base=None
for n in (1,2,4,8,'strong16M_8'):
    f='gpurun_out/r3c_scale_%s.json'%n if isinstance(n,int) else 'gpurun_out/r3c_%s.json'%n
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if n==1: base=d['value']
        print(n, '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], 'eff %.3f'%(d['value']/(base*n)) if isinstance(n,int) and base else '', (d.get('strip_parity') or {}).get('status'), (d.get('strip_kept_block_lists') or {}).get('status'), d.get('strip_exchange'), d['block_list_policy']['since_upload'])
    except Exception as e:
        print(n,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
