mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_draft.py tests/test_gpu_parity_scale.py -x -q -m gpu 2>&1 | tail -4
bash scripts/gpu_profile_r2c.sh
