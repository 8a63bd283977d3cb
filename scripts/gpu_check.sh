# One-call GPU check used during development (run under gpurun): parity tests, smoke, bench.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
cat gpurun_out/bench_latest.json; tail -3 gpurun_out/bench_latest.err
