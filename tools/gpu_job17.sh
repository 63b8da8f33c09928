#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:packed_share_tc -s 3 -c 1 -f -o gpurun_out/prof_k2_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_k2_tc.log 2>&1
