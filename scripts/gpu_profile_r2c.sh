#!/bin/bash
# Final-kernel recapture (after the FindConsensus / traceback / aligner changes): launch list of the default bench command,
# `--set full` captures of the Draft Stage kernels, racecheck on one lane.  Fill / score captures: gpu_profile_r2b.sh.
CMD="python bench.py --config 2 --lanes 1 --contexts 1 --steps 1 --warmup 0 --no-cpu-baseline --other-configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --other-configs= > gpurun_out/r2b_launches_bench.out 2>&1
for k in poa_consensus poa_traceback; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -c 1 -f -o gpurun_out/r2b_full_$k $CMD > gpurun_out/r2b_full_$k.out 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:poa_align_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2b_full_poa_align $CMD > gpurun_out/r2b_full_poa_align.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_align_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r2b_full_poa_map $CMD > gpurun_out/r2b_full_poa_map.out 2>&1
CCS_B200_LANES=1 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_draft.py -m gpu -q -x -k "oracle_small or golden" > gpurun_out/r2b_racecheck_1lane.txt 2>&1; echo "racecheck rc $?" >> gpurun_out/r2b_racecheck_1lane.txt
tail -4 gpurun_out/r2b_racecheck_1lane.txt
