set -x
mkdir -p gpurun_out
for C in 0-3 0-15; do
taskset -c $C python bench.py --no-cpu-baseline --steps 6 > gpurun_out/qc_$C.json 2> gpurun_out/qc_$C.err
python - <<PY
import json
d=json.load(open('gpurun_out/qc_$C.json')); print('cores $C e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
done
