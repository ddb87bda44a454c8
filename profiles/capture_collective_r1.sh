# Evidence for the collective-motion kernels, run on the B200 box through gpurun.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/collective_launches_r1h.csv python profiles/collective_bench.py --cpu-agents 2000 > /dev/null 2> gpurun_out/collective_ncu1.err
ncu --set full --clock-control none --import-source on -k regex:k_herding -s 2 -c 1 -o gpurun_out/prof_herding_r1h -f python profiles/collective_bench.py --cpu-agents 2000 > /dev/null 2> gpurun_out/collective_ncu2.err
ls -la gpurun_out/prof_herding_r1h.ncu-rep
