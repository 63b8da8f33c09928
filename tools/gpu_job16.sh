#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh 20 sda_b200/libsda_b200_q2.so sda_b200/libsda_b200.so sda_b200/libsda_b200_base.so
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err
