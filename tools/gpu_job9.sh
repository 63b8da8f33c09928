#!/bin/bash
# ncu captures of K2 variants (developer job)
mkdir -p gpurun_out
for v in acc2 acc1; do
SDA_B200_LIB=$PWD/sda_b200/libsda_b200_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:packed_share_tc -s 3 -c 1 -f -o gpurun_out/prof_k2_$v \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_k2_$v.log 2>&1
done
