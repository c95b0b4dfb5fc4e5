import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import learn_path_tracing_b200 as L
from oracle import ptoracle as O
from helpers import *
import test_gpu_legacy as T
ctx = L.default_context()
out = {}
for name in ["synthetic", "yoimiya_ground_small"]:
    if name == "synthetic":
        world, cam = synthetic_legacy_world(); W,H=192,128; cam.resolution=(W,H); tree=False
    else:
        world = cached_world(name); W,H=240,160; cam = mesh_camera((W,H)); tree=True
    osc = O.scene_from_legacy_world(world, use_stored_tree=tree)
    rays = T._rays_with_bounces(O, osc, cam, W, H)
    oid, ot = osc.trace(rays)
    gid, gt = world.hit(rays, ctx)
    tris, off = all_triangles(world)
    bad = np.flatnonzero(gid != oid)
    print(name, "rays", len(rays), "mismatch", len(bad), "gpu-miss/oracle-hit", int(((gid<0)&(oid>=0)).sum()), "gpu-hit/oracle-miss", int(((gid>=0)&(oid<0)).sum()))
    g_tri = np.where(gid[bad] >= off, gid[bad]-off, -1).astype(np.int32); o_tri = np.where(oid[bad] >= off, oid[bad]-off, -1).astype(np.int32)
    tg, wg = O.triangle_eval(tris, g_tri, rays[bad]); to, wo = O.triangle_eval(tris, o_tri, rays[bad])
    tie = (gid[bad]>=0)&(oid[bad]>=0)&(np.abs(gt[bad]-ot[bad]) <= 1e-5*np.abs(ot[bad])+1e-7)
    edge = ((g_tri>=0)&(np.abs(wg)<2e-4))|((o_tri>=0)&(np.abs(wo)<2e-4))
    un = ~(tie|edge)
    print("  ties", int(tie.sum()), "edge", int(edge.sum()), "unexplained", int(un.sum()))
    for k in np.flatnonzero(un)[:12]:
        b = bad[k]
        print(f"   ray {b} (secondary={b>=W*H}) gid {gid[b]} gt {gt[b]:.6f} wmin_g {wg[k]:.2e} t_ref(g) {tg[k]:.6f} | oid {oid[b]} ot {ot[b]:.6f} wmin_o {wo[k]:.2e} o={rays[b,:3]} d={rays[b,4:7]}")
    agree = (gid==oid)&(gid>=0)
    rel = np.abs(gt[agree]-ot[agree])/ot[agree]
    print("  t rel err on agreeing hits: max", rel.max(), "p99.9", np.quantile(rel,0.999), "count>1e-5", int((rel>1e-5).sum()))
    w = np.flatnonzero(agree)[rel>1e-5][:8]
    for b in w: print(f"   ray {b} sec={b>=W*H} id {gid[b]} gt {gt[b]:.7f} ot {ot[b]:.7f}")
    # conditioning of the unexplained cases
    def geo(tri_ids, rr):
        T = tris[tri_ids]; p1,p2,p3 = T[:,0:3],T[:,3:6],T[:,6:9]
        N = np.cross(p2-p1,p3-p1); area = np.linalg.norm(N,axis=1); N = N/area[:,None]
        dn = (rr[:,4:7]*N).sum(1); dist = ((p1-rr[:,:3])*N).sum(1)
        return dn, dist, area, np.linalg.norm(p2-p1,axis=1), np.linalg.norm(p3-p1,axis=1)
    k = np.flatnonzero(un)[:12]
    if len(k):
        dn, dist, area, l1, l2 = geo(g_tri[k], rays[bad[k]])
        for i in range(len(k)): print(f"   gpu-tri: d.N {dn[i]:.3e} plane-dist {dist[i]:.3e} area {area[i]:.3e} edges {l1[i]:.3e} {l2[i]:.3e}")
