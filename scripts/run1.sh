set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_pairs.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2a_three.json 2> gpurun_out/r2a_three.err; tail -c 1500 gpurun_out/r2a_three.json
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --model circular > gpurun_out/r2a_circ.json 2> gpurun_out/r2a_circ.err; tail -c 1500 gpurun_out/r2a_circ.json
