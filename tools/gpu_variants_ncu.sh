#!/bin/bash
# GPU job: elapsed cycles, instruction count and pipe utilisation of the share-gen kernel for several library builds, from ncu
# counters (independent of the power-capped SM clock, unlike event timings).  usage: bash tools/gpu_variants_ncu.sh tag name1 name2 ...
tag=$1; shift
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
for v in "$@"; do
  lib=sda_b200/variants/lib_$v.so; [ "$v" = base ] && lib=sda_b200/libsda_b200.so
  SDA_B200_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:${RX:-packed_share_tc} -s 3 -c 1 --csv --log-file gpurun_out/${tag}_${v}.csv \
    python bench.py --steps 1 --warmup 3 --rounds ${ROUNDS:-20} --packed-path ${PPATH:-auto} --no-e2e --no-cpu-baseline --no-round-sweep --no-configs45 > gpurun_out/${tag}_${v}.log 2>&1
  python - <<PY
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/${tag}_${v}.csv")) if len(r) > 10]
    hdr = rows[0]; vals = {}
    for r in rows[1:]:
        d = dict(zip(hdr, r)); vals[d["Metric Name"]] = d["Metric Value"]
    g = lambda k: float(vals[k].replace(",", ""))
    print("${v}: cycles %.3fM  inst %.3fG  issue %.1f%%  alu %.1f%%  fmaheavy %.1f%%  time %.3f ms" % (
        g("sm__cycles_elapsed.max") / 1e6, g("smsp__inst_executed.sum") / 1e9, g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        g("gpu__time_duration.sum") / (1e6 if g("gpu__time_duration.sum") > 1e5 else 1)))
except Exception as e:
    print("${v}: failed", e)
PY
done
