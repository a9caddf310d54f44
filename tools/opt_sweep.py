"""Sweep of the library's tuning options on the bench workload (whole registrations under the reference defaults).
usage: python tools/opt_sweep.py <points> [key=value[,key=value...]] ...   (no settings: one knob at a time around the defaults)
Every setting must give the same iteration count and final RMSE: the options change how the search runs, not what it finds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
src, tgt = synth.make_pair(m, 3, "primary")
if len(sys.argv) > 2:
    SWEEP = [("default", [])] + [(a, [kv.split("=") for kv in a.split(",")]) for a in sys.argv[2:]]
else:
    SWEEP = [("default", [])]
    for key, vals in (("range_max", (32, 96, 128, 192)), ("base_occupancy", (2, 4, 8, 16, 32)), ("walk_bias", (-1, 1)),
                      ("keep_enter", (0.1, 0.3, 0.5)), ("keep_exit", (0.5, 1.0)), ("grid_levels", (2, 3, 4)),
                      ("keep_alpha", (1.5, 3.0)), ("keep_bias", (-1, 1)), ("search_leaf", (2, 8, 16)), ("grid_shift", (-1, 1)),
                      ("grid_coarse", (0, 2))):
        SWEEP += [(f"{key}={v}", [(key, v)]) for v in vals]
for name, settings in SWEEP:
    h = Handle(0); h.set_params(ICPParameters())
    for key, val in settings: h.set_option(key, float(val))
    best = None
    for rep in range(3):
        res = h.register(src.copy(), tgt)
        t = res.timings_ms
        if best is None or t["loop"] < best["loop"]: best = dict(t)
    print(f"{name:34s} it={res.totalIterations:3d} rmse={res.finalRMSE:.12g} loop={best['loop']:7.3f} nn={best['nn_total']:7.3f} first={best['nn_first']:6.3f}", flush=True)
    h.close()
