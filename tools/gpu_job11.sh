#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_bench.py --only packed_reconstruct > gpurun_out/kernels_reveal.jsonl 2> gpurun_out/kernels_reveal.err
SDA_B200_LIB=$PWD/sda_b200/libsda_b200_base.so timeout 600 python tools/kernel_bench.py --only packed_reconstruct > gpurun_out/kernels_reveal_base.jsonl 2> gpurun_out/kernels_reveal_base.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
