mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_draft.py::test_small_device_budget_splits_into_waves" "tests/test_gpu_draft.py::test_draft_edge_cases" -x -q -m gpu > gpurun_out/sanitizer1.log 2>&1
grep -E "Invalid|at 0x|by thread|Address|passed|failed|ERROR SUMMARY|kernel" gpurun_out/sanitizer1.log | head -30
