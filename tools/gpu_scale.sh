#!/bin/bash
# 1/2/4/8-GPU weak-scaling lines on ONE box (gpurun --gpus 8): same command the driver uses
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-round-sweep --no-e2e > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
    bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-round-sweep --no-e2e > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
