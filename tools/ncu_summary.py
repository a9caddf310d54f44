#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (and optionally the `--page source --csv` export) into the handful
of numbers DESIGN.md / profiles/ quote.  usage: ncu_summary.py raw.csv [source.csv] [top_n]"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
for r in data:
    print("==", r[ki] if ki is not None else "")
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = [(float(r[i]), h.split("stalled_")[1].split("_per_")[0]) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
    st.sort(reverse=True)
    print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
if len(sys.argv) > 2:
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(sys.argv[2])))
    # find the header row of the source table
    h = None
    for i, r in enumerate(rows):
        if "Source" in r and any("Sampl" in c for c in r):
            h = i
            break
    if h is None:
        sys.exit(0)
    hdr = rows[h]
    si = hdr.index("Source")
    samp = [i for i, c in enumerate(hdr) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Samples")]
    instc = [i for i, c in enumerate(hdr) if c.strip() in ("Instructions Executed", "# Instructions Executed")]
    print("source columns:", hdr[:12])
    if samp:
        col = samp[0]
        body = [r for r in rows[h + 1:] if len(r) > col and r[col].replace('.', '', 1).isdigit()]
        tot = sum(float(r[col]) for r in body) or 1.0
        body.sort(key=lambda r: -float(r[col]))
        for r in body[:top]:
            ie = r[instc[0]] if instc else ""
            print(f"  {100 * float(r[col]) / tot:5.1f}%  inst={ie:>10s}  {r[si][:110]}")
