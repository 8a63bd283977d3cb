# other BASELINE.json configs through the same bench (under gpurun)
set -x
mkdir -p gpurun_out
python bench.py --config 3 --zmws 600 --steps 4 --warmup 1 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err
python bench.py --config 5 --zmws 600 --steps 4 --warmup 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -2 gpurun_out/bench_cfg5.err
python - <<PY
import json
for c in (3,5):
    d=json.load(open('gpurun_out/bench_cfg%d.json'%c)); print('config',c,'e2e',round(d['e2e']['value'],1),'hifi frac',round(d['hifi_fraction'],3),'cpu',d.get('cpu_baseline'),'roof',round(d['roofline']['frac'],3),d['e2e']['step_s'])
PY
python bench.py --no-cpu-baseline --steps 12 > gpurun_out/bench_k12.json 2> gpurun_out/bench_k12.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_k12.json')); print('K=12 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
