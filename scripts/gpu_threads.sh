set -x
mkdir -p gpurun_out
for cfg in "2 2" "3 2" "4 1" "2 3"; do
set -- $cfg
taskset -c 0-3 python bench.py --no-cpu-baseline --steps 6 --lanes $1 --contexts $2 > gpurun_out/qc4_$1_$2.json 2> gpurun_out/qc4.err
python - <<PY
import json
d=json.load(open('gpurun_out/qc4_$1_$2.json')); print('4 cores L$1 C$2 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
done
