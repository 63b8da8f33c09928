#!/bin/bash
# GPU job: the share-gen bench line for several builds of the library, interleaved twice (boxes drift by 1-2 %).
# usage (under gpurun): bash tools/gpu_variants.sh tag name1 name2 ...   (names of sda_b200/variants/lib_<name>.so; "base" = the default library)
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    lib=sda_b200/variants/lib_$v.so; [ "$v" = base ] && lib=sda_b200/libsda_b200.so
    SDA_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --rounds ${ROUNDS:-20} --no-e2e --no-cpu-baseline --no-round-sweep --no-configs45 \
      > gpurun_out/${tag}_${v}_${rep}.json 2> gpurun_out/${tag}_${v}_${rep}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}_${rep}.json"))
    print("${v} #${rep}", "ms", round(d["roofline"]["ms_per_launch"], 3), "frac", round(d["roofline"]["frac"], 4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("${v} #${rep} failed", e)
PY
  done
done
