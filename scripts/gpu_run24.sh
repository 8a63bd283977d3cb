mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests -q -m gpu -x > gpurun_out/sanitizer2.log 2>&1
grep -E "Invalid|at |by thread|Address|passed|failed|ERROR SUMMARY|kernel|is out of bounds" gpurun_out/sanitizer2.log | head -40
