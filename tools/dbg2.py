import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clouds, torch
from oracle.binding import Oracle
from iterativeclosestpoint_b200.engine import Handle
orc=Oracle()
tgt=clouds.terrain(20000)
q=clouds.query_sets(tgt)['copies']
want=orc.octree(tgt).find_nearest(q)
def dirty(val):
    x=torch.empty(2_000_000_000,dtype=torch.uint8,device='cuda'); x.fill_(val); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
for val in (0x00,0xFF,0x7F,0x3C):
    for mode in (0,1):
        dirty(val)
        h=Handle(0); h.set_option('nn_mode',mode); h.octree_build(tgt)
        idx,dist,_=h.nn_query(q)
        print('fill %02x mode %d bad'%(val,mode),np.flatnonzero(idx!=want)[:10])
        h.set_option('order_queries',0)
        idx,dist,_=h.nn_query(q)
        print('   unordered bad',np.flatnonzero(idx!=want)[:10])
        h.close()
