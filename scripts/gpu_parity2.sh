mkdir -p gpurun_out
CCS_PARITY_SCALE=2 timeout 1200 python -m pytest tests/test_gpu_parity_scale.py -q -m gpu -k "bench_scale" -v 2>&1 | tail -8 > gpurun_out/r2b_parity_scale2.txt
cat gpurun_out/r2b_parity_scale2.txt
