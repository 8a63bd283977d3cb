set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_draft.py tests/test_gpu_arrow.py -x -q -m gpu 2>&1 | tail -5
for L in 1 2 3 4; do
python bench.py --steps 2 --warmup 1 --lanes $L --no-cpu-baseline > gpurun_out/bench_r1_f_l$L.json 2> gpurun_out/bench_r1_f_l$L.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_f_l$L.json')); print('lanes',$L,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3), d['kernel_ms'])"
tail -3 gpurun_out/bench_r1_f_l$L.err
done
