# quick perf probe (under gpurun): draft tests + single-lane and default bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_draft.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 2 --warmup 1 --contexts 1 --lanes 1 --no-cpu-baseline > gpurun_out/q1.json 2> gpurun_out/q1.err
python - <<PY
import json; d=json.load(open('gpurun_out/q1.json')); print('L1 e2e',round(d['e2e']['value'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
PY
python bench.py --no-cpu-baseline > gpurun_out/q2.json 2> gpurun_out/q2.err
python - <<PY
import json; d=json.load(open('gpurun_out/q2.json')); print('default e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
tail -2 gpurun_out/q1.err gpurun_out/q2.err
