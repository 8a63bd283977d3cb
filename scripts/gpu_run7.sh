set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_draft.py tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err
cat gpurun_out/bench_r1_e.json; tail -5 gpurun_out/bench_r1_e.err
