#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh 20 sda_b200/libsda_b200_acc2b.so sda_b200/libsda_b200_acc1b.so
timeout 900 python bench.py --no-cpu-baseline --no-round-sweep > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
