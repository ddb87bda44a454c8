set -x
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
for m in three_circle circular; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --model $m --e2e-steps 1 > gpurun_out/r2d_${m}.json 2> gpurun_out/r2d_${m}.err; python - <<PY
import json; d=json.load(open('gpurun_out/r2d_${m}.json')); print('$m', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
PY
done
