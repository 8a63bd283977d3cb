# final records: default bench (with cpu_baseline), reference arm, configs 2/5 in the same run
mkdir -p gpurun_out
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 2500 gpurun_out/final_bench.json | head -c 1200; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
cat gpurun_out/final_ref.json | head -c 900; echo
nproc
