# arrow_fill_alpha in one deep launch: config 2, 2000 ZMWs on a single lane (CUDA events of the largest launch)
mkdir -p gpurun_out
python bench.py --config 2 --zmws 2000 --lanes 1 --contexts 1 --steps 1 --warmup 1 --no-cpu-baseline --other-configs '' > gpurun_out/r2b_fill_2000zmw.json 2> gpurun_out/r2b_fill_2000zmw.err
python - <<PY
import json; d=json.load(open('gpurun_out/r2b_fill_2000zmw.json'))
for r in d['roofline_kernels'][:3]: print(r['kernel'], 'largest launch', r['largest_launch'], 'single', round(r['single_lane_all_launches']['frac'],3))
print('e2e', d['e2e']['value'])
PY
tail -2 gpurun_out/r2b_fill_2000zmw.err
