cd /root/repo
# high densities: the deterministic in-cell ordering (k_rank_fix) is quadratic in the cell occupancy -- what does a rebuild cost when the crowd jams?
for rho in 1 2 4 6; do
for K in 1 16; do
python bench.py --model circular --agents 1000000 --density $rho --steps 40 --warmup 6 --rebuild-max $K --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/r2z_dense_${rho}_K$K.json 2>gpurun_out/r2z_dense.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2z_dense_${rho}_K$K.json').read().strip().splitlines()[-1]); print('rho=$rho K=$K', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()}, d['block_list_policy']['since_upload'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2z_dense.err').read()[-800:])
PY
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_dense6_r2z.csv python bench.py --model circular --density 6 --steps 2 --warmup 2 --rebuild-max 1 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
agg={}
rd=csv.reader(l for l in open('gpurun_out/launches_dense6_r2z.csv') if not l.startswith('=='))
h=next(rd); ki,vi,mi=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Name')
for r in rd:
    if len(r)>vi and r[mi]=='gpu__time_duration.sum':
        a=agg.setdefault(r[ki].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print('%-34s %3d launches  %9.1f us each'%(k,c,t/c/1000))
PY
