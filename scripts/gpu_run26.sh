set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_r1_o.json 2> gpurun_out/bench_r1_o.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_o.json')); print('e2e',round(d['e2e']['value'],1),'value',round(d['value'],1),'steps',d['e2e']['step_s'],'roof',round(d['roofline']['frac'],3), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
PY
tail -2 gpurun_out/bench_r1_o.err
