# Round-1 profile capture (run under gpurun): launch list of the bench command + ncu --set full of the hot kernels
set -x
mkdir -p gpurun_out
# (1) launch list of the bench command itself
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
tail -1 gpurun_out/r1_launches_bench.log | cut -c1-200
# (2) full captures of the hot kernels on the same workload, single lane, first (full-population) launches
for K in arrow_fill_alpha arrow_fill_beta arrow_score poa_align; do
ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -c 1 -f -o gpurun_out/r1_full_${K} python bench.py --steps 1 --warmup 0 --lanes 1 --contexts 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
