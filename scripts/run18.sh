cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python bench.py --workload hallway > gpurun_out/r2p_hallway.json 2> gpurun_out/r2p_hallway.err; tail -c 1500 gpurun_out/r2p_hallway.json; tail -5 gpurun_out/r2p_hallway.err
python bench.py > gpurun_out/r2p_three_circle.json 2> gpurun_out/r2p_three_circle.err; tail -c 600 gpurun_out/r2p_three_circle.err
python bench.py --model circular > gpurun_out/r2p_circular.json 2> gpurun_out/r2p_circular.err
for m in three_circle circular; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${m}_r2p.csv python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/launches_${m}_r2p.log 2>&1
done
python - <<'PY'
import json
for m in ('three_circle','circular'):
    d=json.loads(open('gpurun_out/r2p_%s.json'%m).read().strip().splitlines()[-1])
    print(m, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['cpu_baseline']['value'])
    print({k:v['value'] for k,v in d['e2e_variants'].items()})
PY
