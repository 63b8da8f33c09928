#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-round-sweep > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
env | grep -i nccl > gpurun_out/nccl_env.txt
