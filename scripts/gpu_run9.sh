set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_full.csv python bench.py --steps 1 --warmup 1 --zmws 1000 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:poa_align -s 5 -c 1 -o gpurun_out/prof_poa_align_r1 python bench.py --steps 1 --warmup 1 --zmws 300 --lanes 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
