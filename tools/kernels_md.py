#!/usr/bin/env python
"""tools/kernel_bench.py's JSON lines -> the markdown table of profiles/rNN_kernels.md.  usage: kernels_md.py in.jsonl > out.md"""
import json
import sys

rows = []
for ln in open(sys.argv[1]):
    try:
        rows.append(json.loads(ln))
    except ValueError:
        continue
print("| kernel / entry point | ms | calls | elements/s | GB/s (algorithmic) | % of HBM peak | sharing-kernel variant |")
print("|---|---|---|---|---|---|---|")
for d in rows:
    if "skipped" in d:
        print(f"| {d['kernel']} | skipped: {d['skipped']} | | | | | |")
        continue
    l2 = "" if d.get("working_set_vs_l2", "").startswith("larger") else " (L2)"
    print(f"| {d['kernel']} | {d['ms_median']:.3f} | {d.get('calls_per_timing', 1)} | {d['elements_per_s']:.3e} | {d['GBps']:.0f} | "
          f"{100 * d['frac_of_hbm_peak']:.1f}{l2} | {d.get('variant', '')} |")
