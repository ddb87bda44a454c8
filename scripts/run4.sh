set -x
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for m in three_circle circular; do for r in 0 1; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --model $m --refinement $r --e2e-steps 1 > gpurun_out/r2c_${m}_$r.json 2> gpurun_out/r2c_${m}_$r.err; python - <<PY
import json; d=json.load(open('gpurun_out/r2c_${m}_$r.json')); print('$m refine $r', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
PY
done; done
for m in three_circle circular; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(k_pair_eval|k_step)$" --launch-skip 20 --launch-count 2 -o gpurun_out/prof_${m}_r2c -f python bench.py --steps 5 --warmup 12 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --model $m > gpurun_out/ncu_${m}_r2c.log 2>&1
done
ls -la gpurun_out | tail -3
