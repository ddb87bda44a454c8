set -x
cd /root/repo
./scripts/pcie_probe > gpurun_out/pcie_probe_r2.txt 2>&1; cat gpurun_out/pcie_probe_r2.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for m in three_circle circular; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_pair_eval|k_stepIL" --launch-skip 30 --launch-count 3 -o gpurun_out/prof_${m}_r2a -f python bench.py --steps 5 --warmup 12 --no-cpu-baseline --no-fp64-peak --e2e-steps 1 --model $m > gpurun_out/ncu_${m}_r2a.log 2>&1
done
ls -la gpurun_out
