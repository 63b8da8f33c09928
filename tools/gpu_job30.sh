#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for v in "" _s8 _s32 _s64; do
  SDA_B200_LIB=$PWD/sda_b200/libsda_b200$v.so timeout 300 python bench.py --steps 3 --participants 16 --no-cpu-baseline --no-round-sweep > gpurun_out/ab8_${i}${v}.json 2> gpurun_out/ab8_${i}${v}.err
done
done
