set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r1_2gpu.json 2> gpurun_out/bench_r1_2gpu.err
cat gpurun_out/bench_r1_2gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_r1_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>> gpurun_out/bench_r1_2gpu.err | cut -c1-400
