set -x
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/q2.json 2> gpurun_out/q2.err
python - <<PY
import json; d=json.load(open('gpurun_out/q2.json')); print('default e2e',round(d['e2e']['value'],1), d['e2e']['step_s'], 'value', round(d['value'],1), 'rounds', d.get('rounds'))
PY
python bench.py --no-cpu-baseline --lanes 4 --contexts 2 > gpurun_out/q3.json 2> gpurun_out/q3.err
python - <<PY
import json; d=json.load(open('gpurun_out/q3.json')); print('L4C2 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
python bench.py --no-cpu-baseline --lanes 3 --contexts 3 > gpurun_out/q4.json 2> gpurun_out/q4.err
python - <<PY
import json; d=json.load(open('gpurun_out/q4.json')); print('L3C3 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
