"""CPU emulation of nn.cu's seeded search on the oracle's tree dump (debug aid)."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clouds, math
from oracle.binding import Oracle
orc=Oracle()
tgt=clouds.terrain(20000)
q_all=clouds.query_sets(tgt)['copies']
ot=orc.octree(tgt); want=ot.find_nearest(q_all)
d=ot.dump()
N=len(d['depth'])
# children lists from preorder
children=[[] for _ in range(N)]; stack=[]
for i in range(N):
    while stack and d['depth'][stack[-1]]>=d['depth'][i]: stack.pop()
    if stack: children[stack[-1]].append(i)
    stack.append(i)
leafstart=np.zeros(N,dtype=np.int64); pos=0
for i in range(N):
    if d['leaf'][i]: leafstart[i]=pos; pos+=d['count'][i]
# full keys for in-leaf order
root=d['box'][0]
def full_key(p):
    lo=[root[0],root[2],root[4]]; hi=[root[1],root[3],root[5]]; k=0
    for lv in range(20):
        o=0
        for a in range(3):
            mid=(lo[a]+hi[a])/2
            if p[a]>mid: o|=1<<a; lo[a]=mid
            else: hi[a]=mid
        k=(k<<3)|o
    return k
def leaf_points(i):
    idx=list(d['idx'][leafstart[i]:leafstart[i]+d['count'][i]])
    idx.sort(key=lambda j:(full_key(tgt[j]),j))
    return idx
def subtree_first_points(i,k=8):
    out=[]
    def rec(n):
        if len(out)>=k: return
        if d['leaf'][n]:
            for j in leaf_points(n):
                if len(out)<k: out.append(j)
        else:
            for c in children[n]: rec(c)
    rec(i); return out
def box(i):
    b=d['box'][i]; return [b[0],b[2],b[4]],[b[1],b[3],b[5]]
def axis_dist(lo,hi,q): 
    m = (lo-q) if not ((lo-q)<(q-hi)) else (q-hi)
    return m if 0.0<m else 0.0
def md_of(i,q):
    lo,hi=box(i); s=[axis_dist(lo[a],hi[a],q[a]) for a in range(3)]
    return math.sqrt((s[0]*s[0]+s[1]*s[1])+s[2]*s[2])
def d2_of(j,q):
    dx=tgt[j]-q; return (dx[0]*dx[0]+dx[1]*dx[1])+dx[2]*dx[2]
def dfs(start,q,best,track,band):
    S={'best':best,'idx':None,'amb':False,'visited':[]}
    def rec(n):
        S['visited'].append(n)
        if d['leaf'][n]:
            fl=False
            for j in leaf_points(n):
                v=d2_of(j,q)
                if track and S['idx'] is None and band[0]<=v<=band[1]: S['amb']=True
                if v<S['best'] or (fl and v==S['best'] and j<S['idx']):
                    S['best']=v; S['idx']=j; fl=True
            return
        ch=[(md_of(c,q),int(d['key'][c])&7,c) for c in children[n]]
        ch.sort(key=lambda t:(t[0],t[1]))
        for md,o,c in ch:
            m=md*md
            if track and S['idx'] is None and band[0]<=m<=band[1]: S['amb']=True
            if m>=S['best']: break
            rec(c)
    rec(start); return S
def seeded(q,verbose=False):
    n=0; path=[]
    while True:
        lo,hi=box(n)
        c=min(min(q[a]-lo[a],hi[a]-q[a]) for a in range(3))
        cf=float(np.nextafter(np.float32(c),np.float32(-np.inf))) if c>0 and float(np.float32(c))>c else (float(np.float32(c)) if c>0 else 0.0)
        path.append((n,cf))
        if d['leaf'][n]: break
        lo_,hi_=box(n); o=0
        for a in range(3):
            if q[a]>(lo_[a]+hi_[a])/2: o|=1<<a
        nxt=[c for c in children[n] if (int(d['key'][c])&7)==o]
        if not nxt: break
        n=nxt[0]
    seeds=subtree_first_points(n,8)
    Sd=min(d2_of(j,q) for j in seeds)
    eps=2.0**-40; hi=Sd*(1+eps)+1e-300; hi2=hi*(1+eps); hi3=hi2*(1+eps)
    start=0
    for l in range(len(path)-1,0,-1):
        if path[l][1]*path[l][1]>hi3: start=path[l][0]; break
    S=dfs(start,q,hi,True,(hi,hi2))
    if verbose: print(' path',path,'seeds',seeds,'Sd',Sd,'start',start,'depth',d['depth'][start],'res',S['idx'],S['best'],'amb',S['amb'],'visited',S['visited'])
    return S
for i in (3794,5389):
    print('query',i,'want',want[i]); seeded(q_all[i],True)
bad=0
for i in range(0,len(q_all)):
    S=seeded(q_all[i])
    if S['idx']!=want[i] and not S['amb']: bad+=1; print('EMU MISMATCH',i,S['idx'],want[i])
print('emu mismatches',bad)
