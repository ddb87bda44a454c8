# Round 2, last GPU call: full suite on the default library, bench with the two-crowd e2e leg, the TMA-staged and the gated
# sweep (CROWD_B200_SWEEP = staged | gated; alt build with 160 staged records per column and one more CTA per SM), ncu captures
# of the new kernels, then the full suite again with the faster sweep forced on.  Ordered by priority: the budget may cut the tail.
cd /root/repo
mkdir -p gpurun_out
T=r4a
ALT=/root/repo/crowddynamics_b200/csrc/alt/libcrowd_b200_cap160.so
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
echo "== 1. full GPU suite, default library"
timeout 600 python -m pytest tests -q -m gpu 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED|^ERROR" | cut -c1-300 | head -40
echo "== 2. bench, default (driver invocation)"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; summ gpurun_out/${T}_bench_default.json
echo "== 3. plain vs staged vs gated sweep, 100 steps"
for ms in three_circle:plain three_circle:staged three_circle:gated circular:plain circular:gated; do
  m=${ms%%:*}; s=${ms##*:}
  CROWD_B200_SWEEP=$s timeout 200 python bench.py --model $m $B > gpurun_out/${T}_${m}_${s}.json 2> gpurun_out/${T}_${m}_${s}.err; summ gpurun_out/${T}_${m}_${s}.json
done
BEST=$(python - <<'PY'
import json
best, bv = 'plain', 0.0
for s in ('plain', 'staged', 'gated'):
    try:
        v = json.loads(open('gpurun_out/r4a_three_circle_%s.json' % s).read().strip().splitlines()[-1])['value']
    except Exception:
        v = 0.0
    if v > bv * (1.01 if s != 'plain' else 1.0):
        best, bv = s, v
print(best)
PY
)
echo "== 6. full GPU suite with the sweep forced to: $BEST"
if [ "$BEST" != "plain" ]; then
  CROWD_B200_SWEEP=$BEST timeout 600 python -m pytest tests -q -m gpu -n 3 2>&1 | grep -E "^E  .*(assert|Error|\{|rror)|passed|failed|^FAILED|^ERROR" | cut -c1-300 | head -30
  CROWD_B200_SWEEP=$BEST timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${BEST}.json 2> gpurun_out/${T}_bench_${BEST}.err; summ gpurun_out/${T}_bench_${BEST}.json
fi
echo "== 5. ncu: k_sweep_gated / k_sweep_staged, three_circle"
NC=$BEST; [ "$BEST" = "plain" ] && NC=gated
for s in $NC; do
  CROWD_B200_SWEEP=$s timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_sweep' -s 10 -c 2 -o gpurun_out/prof_sweep_${s}_${T} -f python bench.py --steps 6 --warmup 8 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1 > gpurun_out/ncu_sweep_${s}_${T}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
echo "== 4. 160 records per column, 7 / 6 CTAs per SM"
for s in staged gated; do
  CROWD_B200_SWEEP=$s CROWD_B200_LIB=$ALT timeout 200 python bench.py --model three_circle $B > gpurun_out/${T}_three_circle_${s}160.json 2> gpurun_out/${T}_three_circle_${s}160.err; summ gpurun_out/${T}_three_circle_${s}160.json
done
echo "== 7. density 0.125"
for s in plain $BEST; do
  CROWD_B200_SWEEP=$s timeout 200 python bench.py --density 0.125 $B > gpurun_out/${T}_rho0125_${s}.json 2> gpurun_out/${T}_rho0125_${s}.err; summ gpurun_out/${T}_rho0125_${s}.json
done
