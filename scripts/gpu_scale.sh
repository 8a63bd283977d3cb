# N-GPU weak-scaling probe (under gpurun --gpus N): bash scripts/gpu_scale.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nproc; nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/bench_scale_$N.json 2> gpurun_out/bench_scale_$N.err
tail -3 gpurun_out/bench_scale_$N.err
python - <<PY
import json; d=json.loads(open("gpurun_out/bench_scale_$N.json").read().strip().splitlines()[-1]); print('gpus',d['n_gpus'],'e2e',round(d['e2e']['value'],1),'per-gpu',round(d['e2e']['value']/d['n_gpus'],1),d['e2e']['step_s'], d['clocks'])
PY
