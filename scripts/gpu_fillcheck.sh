# fill parity tests + single-lane kernel times (run under gpurun)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 2 --warmup 1 --contexts 1 --lanes 1 --no-cpu-baseline > gpurun_out/q1.json 2> gpurun_out/q1.err
python - <<PY
import json; d=json.load(open('gpurun_out/q1.json')); print('L1 e2e',round(d['e2e']['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}); print('roofline', d['roofline'])
PY
tail -n 3 gpurun_out/q1.err
