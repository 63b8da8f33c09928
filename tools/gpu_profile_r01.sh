#!/bin/bash
# round-1 evidence run: tests, bench lines, launch list + full captures of the dominant kernels at the bench configuration
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --packed-path cuda --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/bench_cuda.json 2> gpurun_out/bench_cuda.err
timeout 600 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share_tc -s 3 -c 1 -o gpurun_out/prof_k2_tc \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_k2_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share_m61 -s 3 -c 1 -o gpurun_out/prof_k2_cuda \
    python bench.py --steps 1 --warmup 3 --packed-path cuda --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_k2_cuda.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:combine_kernel -s 30 -c 2 -o gpurun_out/prof_k3 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ncu_k3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:packed_share_combine_tc -s 1 -c 1 -o gpurun_out/prof_fused \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt 2>&1
