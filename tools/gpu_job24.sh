#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for v in "" _nobulk; do
  SDA_B200_LIB=$PWD/sda_b200/libsda_b200$v.so timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-round-sweep > gpurun_out/ab6_${i}${v}.json 2> gpurun_out/ab6_${i}${v}.err
done
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv > gpurun_out/pcie.txt
