# Round-1 evidence capture (run on the B200 box through gpurun): ncu launch lists + full captures of the dominant kernel,
# then the bench lines.  Numbers printed by runs under ncu are never used as bench values.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1f}
for m in three_circle circular; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${m}_r1.csv python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/launches_${m}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -o gpurun_out/prof_${m}_${TAG} -f python bench.py --model $m --steps 2 --warmup 3 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/ncu_${m}.log 2>&1
done
python bench.py > gpurun_out/bench_three_circle.json 2> gpurun_out/bench_three_circle.err
python bench.py --model circular > gpurun_out/bench_circular.json 2> gpurun_out/bench_circular.err
python bench.py --density 0.125 --no-cpu-baseline > gpurun_out/bench_three_circle_rho0125.json 2>/dev/null
python bench.py --model circular --density 0.125 --no-cpu-baseline > gpurun_out/bench_circular_rho0125.json 2>/dev/null
python bench.py --impl reference --steps 5 > gpurun_out/bench_reference.json 2>&1
tail -c 400 gpurun_out/bench_three_circle.json
