#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_bench.py --only mask > gpurun_out/kernels_mask.jsonl 2> gpurun_out/kernels_mask.err
