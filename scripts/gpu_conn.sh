set -x
mkdir -p gpurun_out
for C in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$C python bench.py --no-cpu-baseline > gpurun_out/qconn_$C.json 2> gpurun_out/qconn_$C.err
python - <<PY
import json
d=json.load(open('gpurun_out/qconn_$C.json')); print('conn $C e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
done
CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --no-cpu-baseline --lanes 6 > gpurun_out/qconn_32_6.json 2> gpurun_out/qconn.err
python - <<PY
import json
d=json.load(open('gpurun_out/qconn_32_6.json')); print('conn 32 L6 e2e',round(d['e2e']['value'],1), d['e2e']['step_s'])
PY
