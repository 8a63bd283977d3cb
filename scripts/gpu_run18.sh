set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1_l.json 2> gpurun_out/bench_r1_l.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_l.json')); print('e2e',round(d['e2e']['value'],1),'value',round(d['value'],1),'steps',d['e2e']['step_s'],'roof',round(d['roofline']['frac'],3), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, d['cpu_baseline'], d['clocks'])
PY
tail -2 gpurun_out/bench_r1_l.err
python bench.py --impl reference > gpurun_out/bench_r1_l_ref.json 2>> gpurun_out/bench_r1_l.err; cut -c1-300 gpurun_out/bench_r1_l_ref.json
