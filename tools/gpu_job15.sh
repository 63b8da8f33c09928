#!/bin/bash
mkdir -p gpurun_out
for v in q2; do
SDA_B200_LIB=$PWD/sda_b200/libsda_b200_$v.so timeout 600 python -m pytest tests/test_gpu_device.py tests/test_gpu_parity.py tests/test_gpu_hostpipe.py tests/test_gpu_federated.py -m gpu -q -x > gpurun_out/pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$v.log
done
bash tools/gpu_ab.sh 20 sda_b200/libsda_b200_single.so sda_b200/libsda_b200_pair.so sda_b200/libsda_b200_q2.so
bash tools/gpu_ab.sh 12 sda_b200/libsda_b200_q2.so
bash tools/gpu_ab.sh 8 sda_b200/libsda_b200_q2.so
