import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clouds
from oracle.binding import Oracle
from iterativeclosestpoint_b200.engine import Handle
orc=Oracle()
tgt=clouds.terrain(20000)
q=clouds.query_sets(tgt)['copies']
want=orc.octree(tgt).find_nearest(q)
h=Handle(0); h.set_option('nn_mode',1); h.octree_build(tgt)
h.set_option('dbg_x', float(q[5389,0]))
idx,dist,_=h.nn_query(q)
print('bad',np.flatnonzero(idx!=want)[:10])
h.set_option('order_queries',0)
idx,dist,_=h.nn_query(q)
print('unordered bad',np.flatnonzero(idx!=want)[:10])
