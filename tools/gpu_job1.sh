#!/bin/bash
# first GPU job of the session: parity tests, bench, pipe rates, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./tools/pipe_ubench > gpurun_out/pipe_ubench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r20.json 2> gpurun_out/bench_r20.err
timeout 300 python bench.py --rounds 8 --no-e2e --no-cpu-baseline > gpurun_out/bench_r8.json 2> gpurun_out/bench_r8.err
timeout 300 python bench.py --rounds 12 --no-e2e --no-cpu-baseline > gpurun_out/bench_r12.json 2> gpurun_out/bench_r12.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --participants 64 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share -s 3 -c 1 -o gpurun_out/prof_packed \
    python bench.py --steps 1 --warmup 3 --participants 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_packed.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:combine_kernel -s 15 -c 1 -o gpurun_out/prof_combine \
    python bench.py --steps 1 --warmup 3 --participants 64 --no-e2e --no-cpu-baseline > gpurun_out/ncu_combine.log 2>&1
ls -la gpurun_out
