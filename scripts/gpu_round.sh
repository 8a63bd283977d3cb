# One development round on the GPU box (run under gpurun): parity tests, single-lane kernel times, default bench.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 2 --warmup 1 --contexts 1 --lanes 1 --no-cpu-baseline > gpurun_out/q1.json 2> gpurun_out/q1.err
python - <<PY
import json; d=json.load(open('gpurun_out/q1.json')); print('L1 e2e',round(d['e2e']['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}); print('roofline', d['roofline'])
PY
python bench.py --no-cpu-baseline > gpurun_out/q2.json 2> gpurun_out/q2.err
python - <<PY
import json; d=json.load(open('gpurun_out/q2.json')); print('default e2e',round(d['e2e']['value'],1), d['e2e']['step_s'], 'value', round(d['value'],1))
PY
tail -2 gpurun_out/q1.err gpurun_out/q2.err
