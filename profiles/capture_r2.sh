# Round-2 evidence capture (run on the B200 box through gpurun): ncu launch lists of one bench command per agent model + full
# captures of the three kernels of the once-per-pair pipeline (two consecutive steps, after the rebuild interval has adapted).
# Numbers printed by runs under ncu are never used as bench values.
mkdir -p gpurun_out
TAG=${TAG:-r2t}
for m in three_circle circular; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${m}_${TAG}.csv python bench.py --model $m --steps 12 --warmup 6 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/launches_${m}_${TAG}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_pair_eval|k_finish' -s 24 -c 6 -o gpurun_out/prof_${m}_${TAG} -f python bench.py --model $m --steps 6 --warmup 8 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 > gpurun_out/ncu_${m}_${TAG}.log 2>&1
done
ls -la gpurun_out/*${TAG}*
