import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clouds
from oracle.binding import Oracle
from iterativeclosestpoint_b200.engine import Handle
orc=Oracle()
tgt=clouds.terrain(20000)
q=clouds.query_sets(tgt)['copies']
ot=orc.octree(tgt); want=ot.find_nearest(q)
h0=Handle(0); h0.set_option('nn_mode',0); h0.octree_build(tgt)
for k,qq in clouds.query_sets(tgt).items(): h0.nn_query(qq)
h0.close()
h=Handle(0); h.set_option('nn_mode',1); h.octree_build(tgt)
idx,dist,_=h.nn_query(q)
bad=np.flatnonzero(idx!=want)
print('bad',bad)
d=ot.dump()
# leaf membership
leaf_of=np.empty(len(tgt),dtype=np.int64); pos=0
leaf_nodes=np.flatnonzero(d['leaf']==1)
for ln in leaf_nodes:
    c=d['count'][ln]; leaf_of[d['idx'][pos:pos+c]]=ln; pos+=c
for i in bad:
    d2=((tgt-q[i])**2).sum(1)
    print('q',i,'gpu idx',idx[i],'dist',dist[i],'want',want[i],'d2 gpu pt',d2[idx[i]],'leaf gpu',leaf_of[idx[i]],'leaf want',leaf_of[want[i]],'leafcount',d['count'][leaf_of[want[i]]],'box',d['box'][leaf_of[want[i]]], 'depth', d['depth'][leaf_of[want[i]]])
    # single query alone
    i1,d1,_=h.nn_query(q[i:i+1]); print('  alone:',i1,d1)
