"""CPU prototype used to decide what the GPU builder should do ("Two hierarchies, one sort"):
builds the Yoimiya mesh tree four ways in numpy — Karras LBVH, PLOC (radius 16), top-down binned SAH, top-down
full-sweep SAH, all with one triangle per leaf — and prints the SAH cost in the metric of lbvh.cu:k_sah_cost.
Measured: 14442 / 11728 / 12010 / 11364.  The first two reproduce the GPU builders' costs to all printed digits
(a cross-check of lbvh.cu), and the full-sweep SAH is only 3 % below PLOC (4-7 % fewer node visits when the oracle
walks the trees), which is why no top-down SAH builder was written.  Needs scenes_cache/ (tools/prepare_assets.py).
"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
from helpers import cached_world, mesh_camera
from oracle import ptoracle as O
import learn_path_tracing_b200 as L
sys.setrecursionlimit(100000)

w = cached_world("yoimiya_ground_full")
m = w.meshes[1]
f = m["indices"]; P = m["positions"]
tri9 = P[f[:, [0, 3, 6]]].reshape(-1, 9).astype(np.float32)
v = tri9.reshape(-1, 3, 3)
lo = v.min(1); hi = v.max(1)
pad = 2e-5*np.maximum(np.abs(lo), np.abs(hi)) + 1e-6
lo = (lo - pad).astype(np.float32); hi = (hi + pad).astype(np.float32)
n = len(lo)
cen = 0.5*(lo+hi)

def area(l, h):
    d = h - l
    return d[...,0]*d[...,1] + d[...,1]*d[...,2] + d[...,2]*d[...,0]

class Tree:
    def __init__(self): self.nodes = []  # (lo0,hi0,lo1,hi1,ref0,ref1)
def emit(tree_children, boxes_of):
    pass

# generic: build as nested tuples: leaf = int prim; inner = (left, right)
def to_nodes(root):
    """nested tuple tree -> nodes16 array (root index 0), cost, depth"""
    nodes = []
    def box(t):
        if isinstance(t, (int, np.integer)): return lo[t], hi[t]
        return t[2], t[3]
    # annotate boxes bottom-up iteratively
    def annotate(t):
        if isinstance(t, (int, np.integer)): return t
        l = annotate(t[0]); r = annotate(t[1])
        bl, bh = box(l); cl, ch = box(r)
        return (l, r, np.minimum(bl, cl), np.maximum(bh, ch))
    root = annotate(root)
    out = []
    def alloc(t):
        idx = len(out); out.append(None)
        refs = []
        rec = np.zeros(16, np.float32)
        for k, c in enumerate(t[:2]):
            bl, bh = box(c)
            rec[6*k:6*k+3] = bl; rec[6*k+3:6*k+6] = bh
        out[idx] = rec
        for k, c in enumerate(t[:2]):
            if isinstance(c, (int, np.integer)): r = ~int(c)
            else: r = alloc(c)
            rec[12+k:13+k] = np.array([r], np.int32).view(np.float32)
        return idx
    alloc(root)
    nodes = np.stack(out)
    c0 = area(nodes[:,0:3], nodes[:,3:6]); c1 = area(nodes[:,6:9], nodes[:,9:12])
    return nodes, float(c0.sum()+c1.sum())

def morton_order():
    clo = cen.min(0); ext = cen.max(0) - clo
    q = np.clip((cen - clo)/ext*2097152.0, 0, 2097151).astype(np.uint64)
    def expand(v):
        v = v & np.uint64(0x1fffff)
        v = (v | v << np.uint64(32)) & np.uint64(0x1f00000000ffff)
        v = (v | v << np.uint64(16)) & np.uint64(0x1f0000ff0000ff)
        v = (v | v << np.uint64(8)) & np.uint64(0x100f00f00f00f00f)
        v = (v | v << np.uint64(4)) & np.uint64(0x10c30c30c30c30c3)
        v = (v | v << np.uint64(2)) & np.uint64(0x1249249249249249)
        return v
    keys = expand(q[:,0]) << np.uint64(2) | expand(q[:,1]) << np.uint64(1) | expand(q[:,2])
    order = np.argsort(keys, kind='stable')
    return order, keys[order]

def lbvh():
    order, keys = morton_order()
    keys = [int(k) for k in keys]
    def rec(a, b):  # [a,b] inclusive
        if a == b: return int(order[a])
        ka, kb = keys[a], keys[b]
        if ka == kb: s = (a+b)//2
        else:
            bit = (ka ^ kb).bit_length() - 1
            # first index with that bit set
            lo_, hi_ = a, b
            while lo_ < hi_:
                mid = (lo_+hi_)//2
                if (keys[mid] >> bit) & 1: hi_ = mid
                else: lo_ = mid+1
            s = lo_ - 1
        return (rec(a, s), rec(s+1, b))
    return rec(0, n-1)

def ploc(radius=16):
    order, _ = morton_order()
    cl = [int(i) for i in order]           # tree refs
    L_ = lo[order].copy(); H_ = hi[order].copy()
    while len(cl) > 1:
        c = len(cl)
        best = np.full(c, np.inf); nn = np.full(c, -1)
        idx = np.arange(c)
        for off in list(range(-radius, 0)) + list(range(1, radius+1)):
            j = idx + off
            ok = (j >= 0) & (j < c)
            jj = np.clip(j, 0, c-1)
            a = area(np.minimum(L_, L_[jj]), np.maximum(H_, H_[jj])).astype(np.float32)
            better = ok & ((a < best) | ((a == best) & (jj == (idx ^ 1))))
            best = np.where(better, a, best); nn = np.where(better, jj, nn)
        mutual = nn[nn] == idx
        keep = np.ones(c, bool)
        newcl = list(cl)
        for i in np.flatnonzero(mutual & (idx < nn)):
            j = nn[i]
            newcl[i] = (cl[i], cl[j])
            L_[i] = np.minimum(L_[i], L_[j]); H_[i] = np.maximum(H_[i], H_[j])
            keep[j] = False
        cl = [newcl[i] for i in range(c) if keep[i]]
        L_ = L_[keep]; H_ = H_[keep]
    return cl[0]

def sah_topdown(bins=None):
    """full-sweep SAH (bins=None) or binned SAH; 1 prim per leaf"""
    def rec(ids):
        if len(ids) == 1: return int(ids[0])
        if len(ids) == 2: return (int(ids[0]), int(ids[1]))
        best = (np.inf, None, None)
        for ax in range(3):
            o = ids[np.argsort(cen[ids, ax], kind='stable')]
            l_lo = np.minimum.accumulate(lo[o], 0); l_hi = np.maximum.accumulate(hi[o], 0)
            r_lo = np.minimum.accumulate(lo[o][::-1], 0)[::-1]; r_hi = np.maximum.accumulate(hi[o][::-1], 0)[::-1]
            k = np.arange(1, len(o))
            cost = area(l_lo[:-1], l_hi[:-1])*k + area(r_lo[1:], r_hi[1:])*(len(o)-k)
            if bins and len(o) > bins:
                cand = np.unique((np.arange(1, bins)*len(o))//bins) - 1
                cand = cand[(cand >= 0) & (cand < len(cost))]
                i = cand[np.argmin(cost[cand])]
            else:
                i = int(np.argmin(cost))
            if cost[i] < best[0]: best = (cost[i], o, i+1)
        _, o, s = best
        return (rec(o[:s]), rec(o[s:]))
    return rec(np.arange(n))

def rays_sample():
    cam = mesh_camera((480, 270))
    rays = O.generate_rays(cam.to_struct(), 480, 270, 0, 1) if hasattr(O, 'generate_rays') else None
    return rays

if __name__ == "__main__":
    t0=time.time(); T1 = lbvh(); n1,c1 = to_nodes(T1); print("lbvh cost", c1, time.time()-t0)
    t0=time.time(); T2 = ploc(16); n2,c2 = to_nodes(T2); print("ploc16 cost", c2, time.time()-t0)
    t0=time.time(); T3 = sah_topdown(16); n3,c3 = to_nodes(T3); print("binned16 SAH cost", c3, time.time()-t0)
    t0=time.time(); T4 = sah_topdown(None); n4,c4 = to_nodes(T4); print("sweep SAH cost", c4, time.time()-t0)
    np.savez("/tmp/proto/trees.npz", lbvh=n1, ploc=n2, binned=n3, sweep=n4, tri9=tri9)
