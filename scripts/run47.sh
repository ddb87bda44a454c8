# Round 2, last GPU call: prefetch A/B (k_pair_eval: next pair's records -> L1 / L2; k_finish: contribution lines before the
# plane loads), full suite on the default library, the driver's bench invocation (e2e with 4 crowds in flight), smoke, launch list.
cd /root/repo
mkdir -p gpurun_out
T=r4c
ALT=/root/repo/crowddynamics_b200/csrc/alt
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
echo "== 1. prefetch A/B, three_circle then circular"
for l in default nopf pairL1 fin pairL2_fin; do
  if [ $l = default ]; then unset CROWD_B200_LIB; else export CROWD_B200_LIB=$ALT/lib_$l.so; fi
  timeout 200 python bench.py $B > gpurun_out/${T}_3c_$l.json 2> gpurun_out/${T}_3c_$l.err; summ gpurun_out/${T}_3c_$l.json
done
for l in default nopf; do
  if [ $l = default ]; then unset CROWD_B200_LIB; else export CROWD_B200_LIB=$ALT/lib_$l.so; fi
  timeout 200 python bench.py --model circular $B > gpurun_out/${T}_circ_$l.json 2> gpurun_out/${T}_circ_$l.err; summ gpurun_out/${T}_circ_$l.json
done
unset CROWD_B200_LIB
echo "== 2. full GPU suite, default library"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED|^ERROR" | cut -c1-300 | head -40
echo "== 3. driver invocation"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; summ gpurun_out/${T}_bench_default.json
echo "== 4. smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== 5. launch list of the bench command"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_three_circle_${T}.csv python bench.py --steps 12 --warmup 6 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1 > gpurun_out/launches_three_circle_${T}.log 2>&1
ls -la gpurun_out/launches_three_circle_${T}.csv
