set -x
mkdir -p gpurun_out
for cfgs in "1 4" "2 2" "2 3" "3 2"; do set -- $cfgs
python bench.py --steps 6 --warmup 2 --contexts $1 --lanes $2 --no-cpu-baseline > gpurun_out/bench_r1_m_$1_$2.json 2> gpurun_out/bench_r1_m.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_m_$1_$2.json')); print('ctx',$1,'lanes',$2,'e2e',round(d['e2e']['value'],1),'value',round(d['value'],1),'steps',d['e2e']['step_s'])
PY
tail -2 gpurun_out/bench_r1_m.err
done
