#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for r in 20 12 8; do
timeout 300 python bench.py --rounds $r --no-e2e --no-cpu-baseline > gpurun_out/bench_r$r.json 2> gpurun_out/bench_r$r.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share -s 3 -c 1 -o gpurun_out/prof_packed \
    python bench.py --steps 1 --warmup 3 --participants 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_packed.log 2>&1
