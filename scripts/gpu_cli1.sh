# ccs BAM -> BAM on one GPU: 4 800 config-2 ZMWs written by the simulator, then the command line (windowed polish)
mkdir -p gpurun_out
python - <<'PY'
import ctypes as C, time
from ccs_b200 import sim, simlib
m = sim.synthetic_model(); cfg = sim.get_config(2)
t = time.time()
rc = simlib().ccs_sim_write_subreads_bam(b"/tmp/bench.subreads.bam", b"m64000_000000_000000", m.ctypes.data_as(C.c_void_p), C.byref(cfg), C.c_int64(0), C.c_int32(4800), C.c_int32(1))
print("wrote BAM rc", rc, "in", round(time.time() - t, 1), "s")
PY
ls -la /tmp/bench.subreads.bam*
( time ccs_b200/bin/ccs /tmp/bench.subreads.bam /tmp/out_w.bam --gpus 1 --log-level INFO --batch-size 600 ) 2> gpurun_out/r2b_cli_gpus1.log; tail -12 gpurun_out/r2b_cli_gpus1.log
( time ccs_b200/bin/ccs /tmp/bench.subreads.bam /tmp/out_w0.bam --gpus 1 --log-level INFO --batch-size 600 --window-size 0 ) 2> gpurun_out/r2b_cli_gpus1_nowin.log; tail -6 gpurun_out/r2b_cli_gpus1_nowin.log
nproc
