#!/bin/bash
# GPU job (gpurun --gpus N): the multi-GPU tests, then the bench line at N ranks as the driver launches it.
N=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_c_abi_smoke.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 ${BENCH_FLAGS} > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_n$N.json"))
    print("N=%d value %.4g  ms/step %.3f  frac %.4f  e2e %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
    for k in ("config4_clerk_sum", "config5_e2e", "clerk_combine_x5"):
        if k in d["kernels"]:
            print(" ", k, {kk: (round(vv, 5) if isinstance(vv, float) else vv) for kk, vv in d["kernels"][k].items() if kk != "note"})
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -5 gpurun_out/${tag}_bench_n$N.err
