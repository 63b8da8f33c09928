#!/bin/bash
# GPU job: parity tests + one full ncu capture of the share-gen kernel at the bench configuration.
# usage (under gpurun): bash tools/gpu_prof_k2.sh [tag] [kernel regex] [bench packed path]
tag=${1:-prof}; rx=${2:-packed_share_tc2}; path=${3:-auto}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${rx} -s 3 -c 1 -o gpurun_out/${tag}_k2 \
    python bench.py --steps 1 --warmup 3 --packed-path ${path} --no-e2e --no-cpu-baseline --no-round-sweep --no-configs45 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
