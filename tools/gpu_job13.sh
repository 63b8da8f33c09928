#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err
bash tools/gpu_sanitize.sh
