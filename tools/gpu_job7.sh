#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_device.py -m gpu -q -k "packed_kernels or config3_full" > gpurun_out/pytest_gpu_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_tc.log
for r in 20 8; do
timeout 300 python bench.py --rounds $r --packed-path tc --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/bench_tc_r$r.json 2> gpurun_out/bench_tc_r$r.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share_tc -s 3 -c 1 -o gpurun_out/prof_packed_tc \
    python bench.py --steps 1 --warmup 3 --participants 16 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_packed_tc.log 2>&1
