#!/bin/bash
# GPU job: parity tests, then the share-gen kernel side by side (paired-tile kernel vs first generation), same box.
# usage (under gpurun): bash tools/gpu_ab_k2.sh [tag]
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for path in auto tc1 auto tc1; do
  for r in 20 12; do
    timeout 300 python bench.py --steps 10 --warmup 3 --rounds $r --packed-path $path --no-e2e --no-cpu-baseline --no-round-sweep --no-configs45 \
      > gpurun_out/${tag}_${path}_r${r}.json 2> gpurun_out/${tag}_${path}_r${r}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${path}_r${r}.json"))
    print("${path} r${r}", d["roofline"]["kernel"], "ms", round(d["roofline"]["ms_per_launch"], 3), "frac", round(d["roofline"]["frac"], 4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("${path} r${r} failed", e)
PY
  done
done
