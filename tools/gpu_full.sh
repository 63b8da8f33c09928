#!/bin/bash
# GPU job: parity tests, smoke, the default bench line and the reference arm, as the driver runs them (1 GPU).
tag=${1:-full}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("value %.4g  ms/step %.3f  frac %.4f  e2e %.4g  launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]), d["clocks"])
    for k, v in d["kernels"].items():
        print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("note",)})
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 gpurun_out/${tag}_bench.err
if [ "${REF:-0}" = 1 ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${tag}_ref.json
fi
