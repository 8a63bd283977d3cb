set -x
mkdir -p gpurun_out
for C in 4 8 16 32; do
export CCS_B200_FILL_CPL=$C
timeout 600 python -m pytest tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 2 --warmup 1 --lanes 1 --stage polish --no-cpu-baseline > gpurun_out/bench_r1_h_c$C.json 2> gpurun_out/bench_r1_h_c$C.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_h_c$C.json')); print('CPL',$C,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3),'GB/s',round(d['roofline']['achieved']), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
PY
tail -2 gpurun_out/bench_r1_h_c$C.err
done
