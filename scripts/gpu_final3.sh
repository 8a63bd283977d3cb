# last check of the round: the whole -m gpu suite, smoke, and the Draft Stage profile re-capture on the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_profile_r2c.sh
