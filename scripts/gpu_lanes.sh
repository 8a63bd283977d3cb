set -x
mkdir -p gpurun_out
for cfg in "4 2 16" "4 2 32" "6 2 48" "8 2 64" "3 3 16" "8 1 16"; do
set -- $cfg
python bench.py --no-cpu-baseline --lanes $1 --contexts $2 --host-threads $3 > gpurun_out/ql_$1_$2_$3.json 2> gpurun_out/ql_$1_$2_$3.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ql_$1_$2_$3.json')); print('L$1 C$2 T$3 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
except Exception as e:
    print('L$1 C$2 T$3 failed', open('gpurun_out/ql_$1_$2_$3.err').read()[-600:])
PY
done
