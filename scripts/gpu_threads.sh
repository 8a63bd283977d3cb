set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_draft.py -x -q -m gpu 2>&1 | tail -2
CCS_B200_HOST_PROFILE=1 python bench.py --no-cpu-baseline --lanes 1 --contexts 1 --steps 1 --warmup 1 2>&1 >/dev/null | grep host-profile | tail -20
for C in 0-3 0-15; do
taskset -c $C python bench.py --no-cpu-baseline --steps 6 > gpurun_out/qc_$C.json 2> gpurun_out/qc_$C.err
python - <<PY
import json
d=json.load(open('gpurun_out/qc_$C.json')); print('cores $C e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
done
