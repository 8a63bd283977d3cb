# ncu full captures of the full-population fill launches (single lane) -- run under gpurun
set -x
mkdir -p gpurun_out
for K in arrow_fill_alpha arrow_fill_beta; do
ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -c 1 -o gpurun_out/r1b_full_${K} -f python bench.py --steps 1 --warmup 0 --lanes 1 --contexts 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
