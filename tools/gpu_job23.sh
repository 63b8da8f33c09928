#!/bin/bash
mkdir -p gpurun_out
SDA_B200_LIB=$PWD/sda_b200/libsda_b200_bulk.so timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_bulk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bulk.log
for i in 1 2; do
for v in "" _bulk; do
  SDA_B200_LIB=$PWD/sda_b200/libsda_b200$v.so timeout 300 python bench.py --rounds 20 --packed-path tc --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ab5_${i}${v}.json 2> gpurun_out/ab5_${i}${v}.err
done
done
