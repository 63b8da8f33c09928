#!/bin/bash
# A/B builds of libsda_b200: usage gpu_ab.sh <rounds> lib1 lib2 ...
mkdir -p gpurun_out
r=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  SDA_B200_LIB=$PWD/$lib timeout 300 python bench.py --rounds $r --packed-path tc --no-e2e --no-cpu-baseline --no-round-sweep > gpurun_out/ab_${name}_r$r.json 2> gpurun_out/ab_${name}_r$r.err
done
