# Round 2, confirmation of the final configuration (prefetch for three-circle pairs only): the driver's bench invocation for
# both agent models, then the GPU suite with what is left of the budget.
cd /root/repo
mkdir -p gpurun_out
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r4d_bench_three_circle.json 2> gpurun_out/r4d_bench_three_circle.err; tail -c 600 gpurun_out/r4d_bench_three_circle.err; python -c "
import json; d=json.loads(open('gpurun_out/r4d_bench_three_circle.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('single_crowd_value'))"
timeout 60 python bench.py --model circular --steps 100 --warmup 10 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --e2e-crowds 1 > gpurun_out/r4d_bench_circular.json 2> gpurun_out/r4d_bench_circular.err; python -c "
import json; d=json.loads(open('gpurun_out/r4d_bench_circular.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
