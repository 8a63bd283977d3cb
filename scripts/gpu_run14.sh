set -x
mkdir -p gpurun_out
for C in 4 8; do for L in 3 4; do
export CCS_B200_FILL_CPL=$C
python bench.py --steps 2 --warmup 1 --lanes $L --no-cpu-baseline > gpurun_out/bench_r1_i_c${C}_l$L.json 2> gpurun_out/bench_r1_i.err
python - <<PY
import json; d=json.load(open('gpurun_out/bench_r1_i_c${C}_l$L.json')); print('CPL',$C,'lanes',$L,'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k in ('ms_fill_alpha','ms_fill_beta','ms_score','ms_poa_align')})
PY
done; done
