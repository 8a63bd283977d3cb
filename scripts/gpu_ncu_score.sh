# ncu full capture of the full-population score launch (single lane) -- run under gpurun
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:arrow_score_kernel -c 1 -o gpurun_out/r1b_full_arrow_score -f python bench.py --steps 1 --warmup 0 --lanes 1 --contexts 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
