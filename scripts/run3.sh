set -x
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for m in three_circle circular; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --model $m > gpurun_out/r2b_$m.json 2> gpurun_out/r2b_$m.err; python - <<PY
import json; d=json.load(open('gpurun_out/r2b_$m.json')); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
PY
done
for m in three_circle circular; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_pair_eval|k_step<" --launch-skip 30 --launch-count 3 -o gpurun_out/prof_${m}_r2b -f python bench.py --steps 5 --warmup 12 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --model $m > gpurun_out/ncu_${m}_r2b.log 2>&1
done
ls -la gpurun_out | tail -5
