set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err
cat gpurun_out/bench_r1_d.json; tail -5 gpurun_out/bench_r1_d.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_d_ref.json 2>> gpurun_out/bench_r1_d.err
cat gpurun_out/bench_r1_d_ref.json
