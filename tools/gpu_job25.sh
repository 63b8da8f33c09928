#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for v in "" _rowp; do
  SDA_B200_LIB=$PWD/sda_b200/libsda_b200$v.so timeout 300 python bench.py --rounds 20 --packed-path tc --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ab7_${i}${v}.json 2> gpurun_out/ab7_${i}${v}.err
done
done
