#!/bin/bash
# A/B of K2 tile-pipeline variants (developer job): parity tests on the default build, then bench lines per variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh 20 sda_b200/libsda_b200_base.so sda_b200/libsda_b200_acc2.so sda_b200/libsda_b200_acc1.so sda_b200/libsda_b200_acc1mb6.so
bash tools/gpu_ab.sh 12 sda_b200/libsda_b200_acc2.so sda_b200/libsda_b200_acc1mb6.so
