# full GPU suite + smoke + default bench (no cpu baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 4 --warmup 2 --no-cpu-baseline --other-configs 2,5 > gpurun_out/full1.json 2> gpurun_out/full1.err
python - <<PY
import json; d=json.load(open('gpurun_out/full1.json')); print('e2e',round(d['e2e']['value'],1), [ (k, round(v['e2e'],1), {a:round(b,1) for a,b in v['kernel_ms_single_lane'].items()}) for k,v in d.get('other_configs',{}).items()])
print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
for r in d.get('roofline_kernels',[]): print('  ', r['kernel'], 'timed', round(r['timed_region']['frac'],3), 'single', round(r['single_lane_all_launches']['frac'],3), 'ms', round(r['single_lane_all_launches']['ms'],1), 'largest', round(r['largest_launch']['frac'],3))
print('score items/step', d['score_items_per_step'])
PY
tail -2 gpurun_out/full1.err
