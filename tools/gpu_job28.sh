#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    tools/federated_bench.py --participants 1024 > gpurun_out/fed_8gpu_8192.json 2> gpurun_out/fed_8gpu_8192.err
