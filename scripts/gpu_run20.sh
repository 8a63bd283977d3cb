set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_n.json')); print('e2e',round(d['e2e']['value'],1),'value',round(d['value'],1),'steps',d['e2e']['step_s'],'roof',round(d['roofline']['frac'],3), d['cpu_baseline'], d['clocks'], d['gpu_launches'])
PY
tail -2 gpurun_out/bench_r1_n.err
