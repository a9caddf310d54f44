"""Per-iteration NN / iteration device times of a full registration (profiling aid)."""
import os, sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m=int(sys.argv[1]) if len(sys.argv)>1 else 10_000_000
for regime in (sys.argv[2:] or ['primary','stress']):
    src,tgt=synth.make_pair(m,3,regime)
    for mode in [int(x) for x in os.environ.get("ICP_MODES","6,4").split(",")]:
        h=Handle(0); h.set_option('nn_mode',mode); h.set_option('count', float(os.environ.get('ICP_COUNT','0')))
        for kv in os.environ.get('ICP_OPTS','').split(','):
            if kv: h.set_option(kv.split('=')[0], float(kv.split('=')[1]))
        h.set_params(ICPParameters(maxIterations=int(os.environ.get("ICP_ITERS","16"))))
        w=src.copy(); t0=time.time(); r=h.register(w,tgt); dt=time.time()-t0
        print(f'{regime} m={m} mode={mode} iters={r.loopIterations} wall={dt:.3f}s timings={ {k:round(v,2) for k,v in r.timings_ms.items()} }')
        print('   nn_ms:',[round(i.nnMs,2) for i in r.iterationHistory])
        print('   it_ms:',[round(i.iterMs,2) for i in r.iterationHistory])
        print('   counters', h.nn_counters(), 'tile (per-thread lanes, candidates)', h.nn_tile_counters())
        print('   rmse :',[round(i.rmse,4) for i in r.iterationHistory])
        h.close()
