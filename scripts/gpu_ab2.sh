mkdir -p gpurun_out
for m in 3 4; do
python bench.py --steps 4 --warmup 2 --no-cpu-baseline --other-configs '' --max-poa-reads $m > gpurun_out/ab_poa$m.json 2> gpurun_out/ab_poa$m.err
python - <<PY
import json; d=json.load(open('gpurun_out/ab_poa$m.json')); print('poa reads $m e2e',round(d['e2e']['value'],1), 'hifi', round(d['hifi_fraction'],3), 'items', round(d['score_items_per_step']/1e6,2), {k:round(v,1) for k,v in d['kernel_ms'].items() if k!='note'})
PY
done
