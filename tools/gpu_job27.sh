#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/federated_bench.py --participants 256 > gpurun_out/fed_1gpu_256.json 2> gpurun_out/fed_1gpu_256.err
timeout 900 python tools/federated_bench.py --participants 1024 > gpurun_out/fed_1gpu_1024.json 2> gpurun_out/fed_1gpu_1024.err
