"""Kernel times of the BUILD phase (upload -> first NN) of one warm registration.
  run:    python tools/build_split.py run [m]          (under ncu --metrics gpu__time_duration.sum --csv --log-file L)
  report: python tools/build_split.py report L"""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "run":
    import numpy as np
    from iterativeclosestpoint_b200 import synth
    from iterativeclosestpoint_b200.engine import Handle, ICPParameters
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    src, tgt = synth.make_pair(m, 3, "primary")
    h = Handle(0); h.set_params(ICPParameters(maxIterations=2))
    for _ in range(3):
        res = h.register(src.copy(), tgt)
    print({k: round(float(v), 3) for k, v in res.timings_ms.items()})
else:
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    seq = [(r[ki].split('(')[0].replace('icpb::', '').replace('void ', ''), float(r[vi].replace(',', '')) / 1000) for r in rows[1:]]
    ends = [i for i, (n, _) in enumerate(seq) if n.startswith('stage_b')]
    # the last registration: from after the stage B before its build to its first stat_a
    last_b = ends[-3] if len(ends) >= 3 else -1
    agg = {}
    order = []
    tot = 0.0
    for n, t in seq[last_b + 1:]:
        if n.startswith('stat_a'): break
        if n not in agg: agg[n] = [0.0, 0]; order.append(n)
        agg[n][0] += t; agg[n][1] += 1; tot += t
    for n in order: print(f"{n[:44]:46s} x{agg[n][1]:3d} {agg[n][0]:9.1f} us")
    print(f"{'sum':46s}      {tot:9.1f} us")
