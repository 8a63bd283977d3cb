mkdir -p gpurun_out
run() { # lanes contexts zmws
python bench.py --steps 4 --warmup 2 --no-cpu-baseline --other-configs '' --lanes $1 --contexts $2 --zmws $3 > gpurun_out/m_$1_$2_$3.json 2> gpurun_out/m_$1_$2_$3.err
python - <<PY
import json; d=json.load(open('gpurun_out/m_$1_$2_$3.json')); print('L$1 C$2 Z$3 e2e',round(d['e2e']['value'],1), [round(x,3) for x in d['e2e']['step_s']])
PY
}
run 4 2 600
run 6 2 600
run 8 2 600
run 4 3 600
run 3 3 600
run 6 2 900
run 8 2 1200
run 4 2 900
