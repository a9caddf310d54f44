#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS samples / executed instructions by CUDA source line.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.defaultdict(lambda: [0.0, 0.0, ""]); cur_file = ""; hdr = None; seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        line = int(r[0]); samples = float(r[hdr.index("# Samples")] or 0); inst = float(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    k = (cur_file, line); a = agg[k]; a[0] += samples; a[1] += inst; a[2] = r[1].strip()[:100]
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts:.0f}, total warp-instructions {ti:.0f} (all captured launches)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/ts:5.1f}% smp {100*a[1]/ti:5.1f}% inst  {k[0]}:{k[1]:<4d} {a[2]}")
