#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, no GPU needed) into the handful of numbers the
roofline discussion uses.  usage: ncu_summary.py report.ncu-rep [label]  -> markdown on stdout"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe % (IMAD)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency / inst issued"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"### {label}\n")
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        print(f"kernel `{d['Kernel Name'][0][:110]}`\n")
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in d:
                print(f"| {name} (`{k}`) | {d[k][0]} {d[k][1]} |")
        st = sorted(((float(v[0] or 0), k[len(STALLS):]) for k, v in d.items()
                     if k.startswith(STALLS) and not k.endswith("_not_issued")), reverse=True)
        tot = sum(s for s, _ in st) or 1
        print("\nstall samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in st[:7]) + "\n")


if __name__ == "__main__":
    main()
