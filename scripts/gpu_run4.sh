set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err
cat gpurun_out/bench_r1_c.json; tail -5 gpurun_out/bench_r1_c.err
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread -k regex:arrow_ -c 12 --csv --log-file gpurun_out/inst_r1_c.csv python bench.py --steps 1 --warmup 0 --zmws 400 --no-cpu-baseline > /dev/null 2>&1
