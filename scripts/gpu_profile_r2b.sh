#!/bin/bash
# Round-2 (final kernels: windowed polish, block-anchored POA aligner) profiles: ncu launch list of the default bench command + one `--set full` capture per heavy kernel
# (first = full-population launch of a single-lane config-2 step), SASS listings, sanitizer runs.
# Run through gpurun; reports land in gpurun_out/, summaries are made here by scripts/summarize_profiles.py.
CMD="python bench.py --config 2 --lanes 1 --contexts 1 --steps 1 --warmup 0 --no-cpu-baseline --other-configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --other-configs= > gpurun_out/r2b_launches_bench.out 2>&1
for k in arrow_fill_alpha arrow_fill_beta arrow_score poa_commit poa_consensus poa_traceback; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -c 1 -f -o gpurun_out/r2b_full_$k $CMD > gpurun_out/r2b_full_$k.out 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:poa_align_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2b_full_poa_align $CMD > gpurun_out/r2b_full_poa_align.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:poa_align_kernel --launch-skip 4 -c 1 -f -o gpurun_out/r2b_full_poa_map $CMD > gpurun_out/r2b_full_poa_map.out 2>&1
ls -la gpurun_out/*.ncu-rep
# sanitizers on the small parity tests (memcheck: all kernels; racecheck: shared-memory hazards of the POA / graph / score kernels)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_draft.py tests/test_gpu_arrow.py -m gpu -q -x -k "small or edge or golden or fill or mixed" > gpurun_out/r2b_memcheck.txt 2>&1; echo "memcheck rc $?" >> gpurun_out/r2b_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_draft.py -m gpu -q -x -k "small or edge" > gpurun_out/r2b_racecheck.txt 2>&1; echo "racecheck rc $?" >> gpurun_out/r2b_racecheck.txt
tail -5 gpurun_out/r2b_memcheck.txt gpurun_out/r2b_racecheck.txt
