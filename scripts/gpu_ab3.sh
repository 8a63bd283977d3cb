mkdir -p gpurun_out
for v in 0 10 12; do
CCS_B200_FILL_VARIANT=$v python bench.py --steps 4 --warmup 2 --no-cpu-baseline --other-configs 2 > gpurun_out/ab_fill$v.json 2> gpurun_out/ab_fill$v.err
python - <<PY
import json; d=json.load(open('gpurun_out/ab_fill$v.json')); k=d['kernel_ms']
print('fill variant $v e2e',round(d['e2e']['value'],1), 'alpha',round(k['ms_fill_alpha'],2),'beta',round(k['ms_fill_beta'],2), [ (r['kernel'][:16], round(r['largest_launch']['frac'],3), round(r['timed_region']['frac'],3)) for r in d['roofline_kernels'][:2]], 'c2', round(d['other_configs']['config2']['e2e'],1), round(d['other_configs']['config2']['fill_alpha_largest_launch_GBps']/6548.5,3), {a:round(b,2) for a,b in d['other_configs']['config2']['kernel_ms_single_lane'].items() if 'fill' in a})
PY
done
