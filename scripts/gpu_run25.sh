mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest "tests/test_gpu_draft.py::test_draft_matches_oracle_small" "tests/test_gpu_arrow.py::test_polish_matches_oracle" -q -m gpu -x > gpurun_out/sanitizer_race.log 2>&1
grep -E "Race|hazard|passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/sanitizer_race.log | head -12
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest "tests/test_gpu_draft.py::test_draft_matches_oracle_small" -q -m gpu -x > gpurun_out/sanitizer_sync.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Barrier|divergent" gpurun_out/sanitizer_sync.log | head -8
