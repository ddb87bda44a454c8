set -x
cd /root/repo
for m in three_circle; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(k_finish|k_records|k_cell_count|k_rank_fix|k_scatter)$" --launch-skip 60 --launch-count 5 -o gpurun_out/prof_${m}_r2f -f python bench.py --steps 5 --warmup 12 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --model $m > gpurun_out/ncu_${m}_r2f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 60 --csv --log-file gpurun_out/launches_${m}_r2f.csv python bench.py --steps 30 --warmup 12 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --model $m > /dev/null 2>&1
done
