"""CPU check of the experimental 4-wide tree (learn_path_tracing_b200/csrc/bvh4.h, host-side collapse of the emitted BVH2):
the collapsed tree must be a proper tree over the same leaves, and the oracle walking it must return exactly the hits it
returns on the binary tree it came from.  The binary trees here are built in numpy (median splits) — shape does not matter."""
import numpy as np
import pytest

from conftest import GOLDEN  # noqa: F401  (makes sure conftest's path setup ran)


def _median_bvh2(tris9):
    """BVH2 in the k_refit_emit layout ([N,16]: c0.min c0.max c1.min c1.max, refs) by median splits of the centroids."""
    v = tris9.reshape(-1, 3, 3)
    lo, hi = v.min(1) - 1e-4, v.max(1) + 1e-4
    cen = 0.5 * (lo + hi)
    nodes = []

    def build(ids):
        if len(ids) == 1:
            return ~int(ids[0]), lo[ids[0]], hi[ids[0]]
        ax = int(np.argmax(cen[ids].max(0) - cen[ids].min(0)))
        o = ids[np.argsort(cen[ids, ax], kind="stable")]
        me = len(nodes)
        nodes.append(None)
        r0, l0, h0 = build(o[: len(o) // 2])
        r1, l1, h1 = build(o[len(o) // 2:])
        rec = np.zeros(16, np.float32)
        rec[0:3], rec[3:6], rec[6:9], rec[9:12] = l0, h0, l1, h1
        rec[12:14] = np.array([r0, r1], np.int32).view(np.float32)
        nodes[me] = rec
        return me, np.minimum(l0, l1), np.maximum(h0, h1)

    build(np.arange(len(tris9)))
    return np.stack(nodes)


@pytest.mark.parametrize("n_tri", [2, 3, 5, 1000, 20011])
def test_collapsed_tree_is_proper_and_returns_the_same_hits(oracle, n_tri):
    tris = oracle.random_triangles(n_tri, 4711, 0.05 if n_tri > 100 else 0.4)
    if n_tri > 100:
        tris[n_tri // 2: n_tri // 2 + 50] = tris[7]     # exact duplicates: lowest id must win on both trees
    nodes2 = _median_bvh2(tris)
    assert nodes2.shape[0] == n_tri - 1
    wide = oracle.bvh4_collapse(nodes2)
    refs = wide[:, 24:28].copy().view(np.int32)
    leaves = np.sort(~refs[(refs < 0)])
    assert np.array_equal(leaves, np.arange(n_tri))                     # every primitive is exactly one leaf
    inner = np.sort(refs[(refs >= 0) & (refs != 0x7FFFFFFF)])
    assert np.array_equal(inner, np.arange(1, wide.shape[0]))           # every wide node but the root has one parent
    arity = ((refs != 0x7FFFFFFF).sum(1))
    assert arity.min() >= 2 and wide.shape[0] <= max(1, (n_tri + 1) // 2 + 1)
    # a child box lies inside its slot's box in the parent: check via the root-to-leaf union for a sample of nodes
    k = np.flatnonzero((refs >= 0) & (refs != 0x7FFFFFFF))
    for flat in k[:: max(1, len(k) // 200)]:
        p, s = divmod(int(flat), 4)
        c = refs[p, s]
        cm = refs[c] != 0x7FFFFFFF
        assert wide[c, 0:4][cm].min() >= wide[p, 0 + s] and wide[c, 12:16][cm].max() <= wide[p, 12 + s]
        assert wide[c, 4:8][cm].min() >= wide[p, 4 + s] and wide[c, 16:20][cm].max() <= wide[p, 16 + s]
        assert wide[c, 8:12][cm].min() >= wide[p, 8 + s] and wide[c, 20:24][cm].max() <= wide[p, 20 + s]
    rays = oracle.random_rays(20000, 99)
    i2, t2, c2 = oracle.trace_bvh2(nodes2, tris, rays)
    i4, t4, c4 = oracle.trace_bvh4(wide, tris, rays)
    assert np.array_equal(i2, i4) and np.array_equal(t2, t4)
    if n_tri > 100:
        assert (i2 >= 0).mean() > 0.05
        assert c4[0] < 0.75 * c2[0]          # fewer steps (prototype: 0.5-0.65x)
        bi, bt = oracle.trace_triangles(tris, rays)[:2]
        assert np.array_equal(i4, bi) and np.array_equal(t4, bt)   # and both equal the brute-force loop
