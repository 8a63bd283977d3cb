set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err
cat gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
# launch list (short run)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --zmws 200 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
# full capture of the two hot kernels
ncu --set full --clock-control none --import-source on -k regex:arrow_fill_alpha -s 1 -c 1 -o gpurun_out/prof_fill_alpha_r1 python bench.py --steps 1 --warmup 1 --zmws 200 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:arrow_score -s 1 -c 1 -o gpurun_out/prof_score_r1 python bench.py --steps 1 --warmup 1 --zmws 200 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
