# Round 2, GPU call after run45: distance_three_circles with fp32-ranked part pairs (PAIR_HMIN_RANKED, default on) -- full
# suite, A/B against the nine-fold loop (alt build), e2e with 2 / 3 / 4 crowds in flight, ncu evidence for k_sweep_staged and
# the new k_pair_eval.
cd /root/repo
mkdir -p gpurun_out
T=r4b
ALT=/root/repo/crowddynamics_b200/csrc/alt/libcrowd_b200_hmin9.so
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    print(sys.argv[1], '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], {k:round(v,4) for k,v in (r.get('phase_ms_per_step') or {}).items()},
          'e2e %.4g'%d['e2e']['value'] if d.get('e2e') else '', 'single %.4g'%d['e2e'].get('single_crowd_value',0) if d.get('e2e') else '',
          d['block_list_policy']['since_upload'])
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
B="--steps 100 --warmup 10 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1"
echo "== 1. full GPU suite"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED|^ERROR" | cut -c1-300 | head -40
echo "== 2. A/B: ranked (default) vs nine-fold loop, three_circle 100 steps, twice"
for i in 1 2; do
  timeout 200 python bench.py $B > gpurun_out/${T}_ranked_$i.json 2> gpurun_out/${T}_ranked_$i.err; summ gpurun_out/${T}_ranked_$i.json
  CROWD_B200_LIB=$ALT timeout 200 python bench.py $B > gpurun_out/${T}_hmin9_$i.json 2> gpurun_out/${T}_hmin9_$i.err; summ gpurun_out/${T}_hmin9_$i.json
done
echo "== 3. driver invocation, e2e with 2 / 3 / 4 crowds in flight"
for c in 2 3 4; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-crowds $c > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; summ gpurun_out/${T}_bench_c$c.json
done
echo "== 4. density 0.125 (ranked)"
timeout 200 python bench.py --density 0.125 $B > gpurun_out/${T}_rho0125.json 2> gpurun_out/${T}_rho0125.err; summ gpurun_out/${T}_rho0125.json
echo "== 5. ncu: k_pair_eval (ranked), k_sweep_staged"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_pair_eval' -s 10 -c 2 -o gpurun_out/prof_pair_eval_${T} -f python bench.py --steps 6 --warmup 8 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1 > gpurun_out/ncu_pair_eval_${T}.log 2>&1
CROWD_B200_SWEEP=staged timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_sweep' -s 10 -c 2 -o gpurun_out/prof_sweep_staged_${T} -f python bench.py --steps 6 --warmup 8 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1 > gpurun_out/ncu_sweep_staged_${T}.log 2>&1
ls -la gpurun_out/*${T}*.ncu-rep
