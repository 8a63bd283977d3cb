mkdir -p gpurun_out
run() { # tag envs...
tag=$1; shift
env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline --other-configs '' > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
python - <<PY
import json; d=json.load(open('gpurun_out/ab_$tag.json')); print('$tag e2e',round(d['e2e']['value'],1), [round(x,3) for x in d['e2e']['step_s']], 'score ms', round(d['kernel_ms']['ms_score'],1))
PY
}
run base X=1
run v0 CCS_B200_SCORE_VARIANT=0
run v1 CCS_B200_SCORE_VARIANT=1
run v4 CCS_B200_SCORE_VARIANT=4
run prio0 CCS_B200_PRIO=0
run halo20 CCS_B200_QV_HALO=20
