"""CPU prototype for the next round: collapse the Yoimiya PLOC tree (saved by tools/tree_quality_proto.py into
/tmp/proto/trees.npz) into a 4-wide tree (largest-area inner child opened first) and count ordered, pruned traversal
steps.  Measured: 23424 BVH2 nodes -> 11342 wide nodes of average arity 3.07; secondary rays 34.6 -> 17.8 steps
(0.51x), primary 4.0 -> 2.5 (0.64x), triangle tests unchanged — and 17.8 x 3.07 = 55 box tests instead of 69.
"""
import sys, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
from helpers import mesh_camera
from oracle import ptoracle as O
d = np.load('/tmp/proto/trees.npz'); tri9 = d['tri9']; nodes = d['ploc']
kids = nodes[:,12:14].copy().view(np.int32)
def boxes(i): return [(nodes[i,0:3], nodes[i,3:6], int(kids[i,0])), (nodes[i,6:9], nodes[i,9:12], int(kids[i,1]))]
def area(l,h):
    e=h-l; return e[0]*e[1]+e[1]*e[2]+e[2]*e[0]
# collapse: a wide node = children of node i, repeatedly replacing the inner child of largest area by its two children, up to 4
wide = {}
def collapse(i):
    ch = boxes(i)
    while len(ch) < 4:
        cand = [(area(l,h), k) for k,(l,h,r) in enumerate(ch) if r >= 0]
        if not cand: break
        _, k = max(cand)
        l,h,r = ch.pop(k); ch += boxes(r)
    wide[i] = ch
    for l,h,r in ch:
        if r >= 0: collapse(r)
sys.setrecursionlimit(100000); collapse(0)
print("bvh2 inner nodes", len(nodes), "bvh4 nodes", len(wide), "avg arity", np.mean([len(c) for c in wide.values()]))
v = tri9.reshape(-1,3,3)
def tri_hit(o, dd, p):
    v0=v[p,0]; e1=v[p,1]-v0; e2=v[p,2]-v0
    pv=np.cross(dd,e2); det=e1@pv
    if det==0: return np.inf
    inv=1/det; tv=o-v0; u=(tv@pv)*inv; qv=np.cross(tv,e1); w=(dd@qv)*inv; t=(e2@qv)*inv
    return t if (u>0 and w>0 and 1-u-w>0 and t>1e-4) else np.inf
def trav(o, dd, wide_mode):
    inv = 1.0/np.where(np.abs(dd)<1e-18,1e-18,dd); best=np.inf; steps=0; tests=0; st=[0]
    while st:
        cur=st.pop()
        if cur<0:
            tests+=1; best=min(best,tri_hit(o,dd,~cur)); continue
        steps+=1
        hits=[]
        for l,h,r in (wide[cur] if wide_mode else boxes(cur)):
            t0=(l-o)*inv; t1=(h-o)*inv
            tn=max(np.minimum(t0,t1).max(),0.0); tf=min(np.maximum(t0,t1).min(),best)
            if tn<=tf: hits.append((tn,r))
        hits.sort(key=lambda x:-x[0])
        for _,r in hits: st.append(r)
    return steps,tests
cam = mesh_camera((480,270)); rays = O.generate_rays(cam.to_struct(),480,270,0,1)
ids,t,_ = O.trace_bvh2(nodes,tri9,rays); hit=np.flatnonzero(ids>=0)
rng=np.random.default_rng(1); sel=rng.choice(hit,1500,replace=False)
o=rays[sel,:3]+t[sel,None]*rays[sel,4:7]; dd=rng.normal(size=o.shape); dd/=np.linalg.norm(dd,axis=1,keepdims=True)
prim=rays[rng.choice(len(rays),3000,replace=False)]
for mode,name in [(False,"bvh2"),(True,"bvh4")]:
    a=np.mean([trav(r[:3].astype(float),r[4:7].astype(float),mode) for r in prim],axis=0)
    b=np.mean([trav(o[i].astype(float),dd[i],mode) for i in range(len(o))],axis=0)
    print(name,"primary steps/tests",a.round(2),"secondary",b.round(2))
