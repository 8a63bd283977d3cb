mkdir -p gpurun_out
run() { # lanes contexts zmws
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --other-configs '' --lanes $1 --contexts $2 --zmws $3 > gpurun_out/m_$1_$2_$3.json 2> gpurun_out/m_$1_$2_$3.err
python - <<PY
import json; d=json.load(open('gpurun_out/m_$1_$2_$3.json')); print('L$1 C$2 Z$3 e2e',round(d['e2e']['value'],1), [round(x,3) for x in d['e2e']['step_s']])
PY
}
run 4 2 600
run 4 3 400
run 4 3 450
run 4 4 300
run 3 3 400
run 4 2 400
