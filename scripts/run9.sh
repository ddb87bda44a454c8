cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for v in A P; do
  if [ $v = A ]; then unset CROWD_B200_LIB; else export CROWD_B200_LIB=/root/repo/crowddynamics_b200/csrc/variants/lib_$v.so; fi
  for m in three_circle circular; do
    timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --model $m --e2e-steps 1 --no-fp64-peak > gpurun_out/r2h_${m}_$v.json 2> gpurun_out/r2h_${m}_$v.err; python - <<PY
import json; d=json.load(open('gpurun_out/r2h_${m}_$v.json')); print('$v $m', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
PY
  done
done
