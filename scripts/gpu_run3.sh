set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
cat gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
