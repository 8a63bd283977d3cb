#!/bin/bash
# 8-GPU session: the scaling lines of BASELINE.json configs 3 (default), 5 (weak + strong), 4 (time-boxed) and the
# BAM -> BAM command line at 1 and 8 GPUs.  Run through gpurun --gpus 8; everything lands in gpurun_out/.
set -x
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
nproc > gpurun_out/scale${N}_nproc.txt; nvidia-smi -L >> gpurun_out/scale${N}_nproc.txt
run --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/scale${N}_c3.json 2> gpurun_out/scale${N}_c3.err
run --config 2 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/scale${N}_c2.json 2> gpurun_out/scale${N}_c2.err
run --config 5 --steps 4 --warmup 1 --no-cpu-baseline > gpurun_out/scale${N}_c5_weak.json 2> gpurun_out/scale${N}_c5_weak.err
run --config 5 --zmws $((800 / N)) --steps 4 --warmup 1 --no-cpu-baseline > gpurun_out/scale${N}_c5_strong.json 2> gpurun_out/scale${N}_c5_strong.err
run --config 4 --minutes 0.5 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/scale${N}_c4_timeboxed.json 2> gpurun_out/scale${N}_c4.err
for f in c3 c2 c5_weak c5_strong c4_timeboxed; do echo $f; python -c "
import json,sys
d=json.loads(open('gpurun_out/scale${N}_$f.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d['steps'], d['config']['zmws_per_step_per_gpu'], d['clocks'])
"; done
# BAM -> BAM through the command line
python - <<'PY'
import ctypes as C, time
from ccs_b200 import sim, simlib
m = sim.synthetic_model(); cfg = sim.get_config(2)
t = time.time()
rc = simlib().ccs_sim_write_subreads_bam(b"/tmp/bench.subreads.bam", b"m64000_000000_000000", m.ctypes.data_as(C.c_void_p), C.byref(cfg), C.c_int64(0), C.c_int32(2400), C.c_int32(1))
print("wrote BAM rc", rc, "in", round(time.time() - t, 1), "s")
PY
ls -la /tmp/bench.subreads.bam*
for g in 1 $N; do ( time ccs_b200/bin/ccs /tmp/bench.subreads.bam /tmp/out_g$g.bam --gpus $g --log-level INFO ) 2> gpurun_out/cli_gpus$g.log; tail -12 gpurun_out/cli_gpus$g.log; done
cmp /tmp/out_g1.bam /tmp/out_g$N.bam && echo "CLI output identical for 1 and $N GPUs"
