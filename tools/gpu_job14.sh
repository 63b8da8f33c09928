#!/bin/bash
mkdir -p gpurun_out
for v in kc kcpair; do
SDA_B200_LIB=$PWD/sda_b200/libsda_b200_$v.so timeout 600 python -m pytest tests/test_gpu_device.py tests/test_gpu_parity.py tests/test_gpu_hostpipe.py -m gpu -q -x > gpurun_out/pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$v.log
done
bash tools/gpu_ab.sh 20 sda_b200/libsda_b200_single.so sda_b200/libsda_b200_pair.so sda_b200/libsda_b200_kc.so sda_b200/libsda_b200_kcpair.so
bash tools/gpu_ab.sh 12 sda_b200/libsda_b200_single.so sda_b200/libsda_b200_kcpair.so
