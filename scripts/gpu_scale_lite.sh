#!/bin/bash
# N-GPU scaling lines of the two main configurations (run through gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
nproc > gpurun_out/r2b_scale${N}_nproc.txt; nvidia-smi -L >> gpurun_out/r2b_scale${N}_nproc.txt
run --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_scale${N}_c3.json 2> gpurun_out/r2b_scale${N}_c3.err
run --config 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_scale${N}_c2.json 2> gpurun_out/r2b_scale${N}_c2.err
for f in c3 c2; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b_scale${N}_$f.json').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d['steps'], d['config']['zmws_per_step_per_gpu'], d['clocks'])
"; done
cat gpurun_out/r2b_scale${N}_nproc.txt | head -1
