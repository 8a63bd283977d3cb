set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 2 --warmup 1 --contexts 1 --lanes 1 --no-cpu-baseline > gpurun_out/bench_r1_p.json 2> gpurun_out/bench_r1_p.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_p.json')); print('e2e',round(d['e2e']['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
PY
tail -2 gpurun_out/bench_r1_p.err
