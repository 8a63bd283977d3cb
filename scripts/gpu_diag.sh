mkdir -p gpurun_out
python scripts/diag_parity.py > gpurun_out/diag_parity.txt 2>&1
tail -30 gpurun_out/diag_parity.txt
