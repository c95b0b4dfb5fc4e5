// extend.cuh — closest-hit queries: World.hit of the reference (10_final/world.py:24-34,
// legacy 15_module.py:838-848) re-designed as ordered, t-pruned BVH2 traversal over a GPU-built LBVH
// plus a short list of "global" primitives that are too large for the tree.
#pragma once
#include "pt_internal.h"
#include "pt_math.cuh"

struct Hit {
    float t;   // -1 on miss
    int prim;  // -1 on miss
    float u, v;  // triangle barycentrics of vertices b and c (Moller-Trumbore); 0 for spheres
};

// Sphere.hit (world.py:43-60) in the REFERENCE'S float32 arithmetic: every operation is an explicitly
// rounded intrinsic (never contracted into FMA), evaluated in the reference's order, so t is bit-identical
// to the CPU oracle.  The reference's b = 2*oc.rd, disc = b*b - 4c, t = (-b -+ sqrt(disc))/2 is written
// in the half-b form, which is exactly equal in binary floating point (scaling by 2 and 4 is exact).
// Returns false when the discriminant is negative (reference leaves t = -1).
PT_DEV bool sphere_hit_ref(float3 o, float3 d, float4 cr, float r2, bool transparent, float tmin, float* t_out) {
    float ocx = __fsub_rn(o.x, cr.x), ocy = __fsub_rn(o.y, cr.y), ocz = __fsub_rn(o.z, cr.z);
    float h = __fadd_rn(__fadd_rn(__fmul_rn(ocx, d.x), __fmul_rn(ocy, d.y)), __fmul_rn(ocz, d.z));
    float c = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(ocx, ocx), __fmul_rn(ocy, ocy)), __fmul_rn(ocz, ocz)), r2);
    float disc = __fsub_rn(__fmul_rn(h, h), c);
    if (!(disc >= 0.0f)) return false;
    // both roots lie behind tmin when -h does (the near root is -h - s <= -h; the far root is only taken by
    // transparent spheres): such a hit is rejected by every caller (t >= tmin), so the square root is skipped
    if (!transparent && -h < tmin) { *t_out = -INFINITY; return true; }
    float s = __fsqrt_rn(disc);
    float t = __fsub_rn(-h, s);
    if (t < PT_EPS && transparent) t = __fadd_rn(-h, s);  // world.py:55-56 far-root rule
    *t_out = t;
    return true;
}

// Moller-Trumbore on precomputed (v0, e1, e2).  Strict inside test like the reference's w1,w2,w3 > 0
// (15_module.py:928); the reference's plane + sub-triangle-area formulation gives the same t to ~1e-6
// relative, edge-grazing rays are the "ties excepted" of the parity contract.
PT_DEV bool triangle_hit_mt(float3 o, float3 d, float3 v0, float3 e1, float3 e2, float* t, float* u, float* v) {
    float3 pvec = cross(d, e2);
    float det = dot(e1, pvec);
    float inv = 1.0f / det;
    float3 tvec = o - v0;
    float uu = dot(tvec, pvec) * inv;
    float3 qvec = cross(tvec, e1);
    float vv = dot(d, qvec) * inv;
    float tt = dot(e2, qvec) * inv;
    *t = tt; *u = uu; *v = vv;
    return uu > 0.0f && vv > 0.0f && (1.0f - uu - vv) > 0.0f;
}

struct TraceCounters {
    unsigned int nodes, prims;
};

template <bool COUNT>
PT_DEV void test_prim(const SceneView& sv, int p, float3 o, float3 d, float tmin, Hit& h, float& best,
                      TraceCounters& tc) {
    if (COUNT) tc.prims++;
    float t, u = 0.0f, v = 0.0f;
    bool ok;
    if (p < sv.n_sph) {
        float4 cr = __ldg(&sv.sph_cr[p]);
        float4 aux = __ldg(&sv.sph_aux[p]);
        ok = sphere_hit_ref(o, d, cr, aux.x, __float_as_int(aux.y) != 0, tmin, &t);
    } else {
        const float4* g = sv.tri_geo + 3 * (size_t)(p - sv.n_sph);
        float4 v0 = __ldg(g), e1 = __ldg(g + 1), e2 = __ldg(g + 2);
        ok = triangle_hit_mt(o, d, f3(v0), f3(e1), f3(e2), &t, &u, &v);
    }
    // closest wins; on exactly equal t the lower primitive id wins (the reference's first-wins order)
    if (ok && t >= tmin && (t < best || (t == best && p < h.prim))) {
        best = t;
        h.t = t; h.prim = p; h.u = u; h.v = v;
    }
}

// "global" triangles (the two halves of a ground plane) from the kernel parameter bank, like the inline spheres: every
// segment of a legacy scene tests them first, so their operands are constant-bank reads instead of an index load and
// three 128-bit loads per lane and segment.  Same arithmetic and tie rule as test_prim.
template <bool COUNT>
PT_DEV void test_inline_tris(const SceneView& sv, float3 o, float3 d, float tmin, Hit& h, float& best, TraceCounters& tc) {
    for (int g = 0; g < sv.n_inl_tri; ++g) {
        if (COUNT) tc.prims++;
        float t, u, v;
        const int p = sv.inl_tri_id[g];
        const bool ok = triangle_hit_mt(o, d, f3(sv.inl_tri[g][0]), f3(sv.inl_tri[g][1]), f3(sv.inl_tri[g][2]), &t, &u, &v);
        if (ok && t >= tmin && (t < best || (t == best && p < h.prim))) {
            best = t;
            h.t = t; h.prim = p; h.u = u; h.v = v;
        }
    }
}

// Ordered traversal with best-t pruning and an explicit stack (LBVH depth is bounded by the 63 Morton
// bits + log2 of duplicate runs; 64 entries cover every tree the builder can emit for < 2^31 prims).
template <bool COUNT>
PT_DEV Hit closest_hit(const SceneView& sv, float3 o, float3 d, float tmin, float tmax, TraceCounters& tc) {
    Hit h;
    h.t = -1.0f; h.prim = -1; h.u = 0.0f; h.v = 0.0f;
    float best = tmax;
    // global spheres straight from the parameter (constant) bank
#pragma unroll
    for (int g = 0; g < PT_MAX_INLINE; ++g) {
        if (g < sv.n_inl) {
            if (COUNT) tc.prims++;
            float t;
            const int p = sv.inl_id[g];
            if (sphere_hit_ref(o, d, sv.inl_cr[g], sv.inl_r2[g], sv.inl_transparent[g] != 0, tmin, &t) && t >= tmin &&
                (t < best || (t == best && p < h.prim))) {
                best = t;
                h.t = t; h.prim = p;
            }
        }
    }
    test_inline_tris<COUNT>(sv, o, d, tmin, h, best, tc);
    for (int g = 0; g < sv.n_global; ++g) test_prim<COUNT>(sv, __ldg(&sv.global_prims[g]), o, d, tmin, h, best, tc);
    if (sv.root == PT_NO_BVH) return h;

    // 1/d with |d| clamped away from zero: keeps lo*inv + oi free of inf - inf = NaN for axis-parallel rays
    const float3 inv = f3(1.0f / (fabsf(d.x) < 1e-18f ? copysignf(1e-18f, d.x) : d.x),
                          1.0f / (fabsf(d.y) < 1e-18f ? copysignf(1e-18f, d.y) : d.y),
                          1.0f / (fabsf(d.z) < 1e-18f ? copysignf(1e-18f, d.z) : d.z));
    const float3 oi = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    int stack[64];
    int sp = 0;
    int cur = sv.root;
    for (;;) {
        if (cur < 0) {
            test_prim<COUNT>(sv, ~cur, o, d, tmin, h, best, tc);
        } else {
            if (COUNT) tc.nodes++;
            const float4* n = sv.nodes + 4 * (size_t)cur;
            const float4 a = __ldg(n), b = __ldg(n + 1), c = __ldg(n + 2), k = __ldg(n + 3);
            // child 0: min (a.x,a.y,a.z) max (a.w,b.x,b.y); child 1: min (b.z,b.w,c.x) max (c.y,c.z,c.w)
            float x0 = fmaf(a.x, inv.x, oi.x), x1 = fmaf(a.w, inv.x, oi.x);
            float y0 = fmaf(a.y, inv.y, oi.y), y1 = fmaf(b.x, inv.y, oi.y);
            float z0 = fmaf(a.z, inv.z, oi.z), z1 = fmaf(b.y, inv.z, oi.z);
            float t0a = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            float t1a = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best));
            x0 = fmaf(b.z, inv.x, oi.x); x1 = fmaf(c.y, inv.x, oi.x);
            y0 = fmaf(b.w, inv.y, oi.y); y1 = fmaf(c.z, inv.y, oi.y);
            z0 = fmaf(c.x, inv.z, oi.z); z1 = fmaf(c.w, inv.z, oi.z);
            float t0b = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            float t1b = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best));
            const bool ha = t0a <= t1a, hb = t0b <= t1b;
            const int ca = __float_as_int(k.x), cb = __float_as_int(k.y);
            if (ha && hb) {
                const bool a_first = t0a <= t0b;
                stack[sp++] = a_first ? cb : ca;
                cur = a_first ? ca : cb;
                continue;
            } else if (ha) {
                cur = ca;
                continue;
            } else if (hb) {
                cur = cb;
                continue;
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return h;
}

// ------------------------------------------------------------------------------------------------
// Resumable traversal for the persistent kernels (persist.cu).  Same result as closest_hit() — closest
// t >= tmin, equal t resolved to the lower primitive id — but cut into single steps (node_step / leaf_step)
// whose state lives in Trav and the per-lane stack, so a warp can decide by ballot WHICH step its lanes take
// next (persist.cu).  A lane that reaches a leaf may wait a few node steps of its neighbours before the leaf
// is tested, i.e. pruning may use a slightly stale `best`; that only costs a few extra node visits, the
// closest hit is order-independent.
#define PT_SENTINEL PT_NO_BVH
#define PT_STACK 64
#ifndef PT_BLOCK
#define PT_BLOCK 256
#endif

// Traversal stack.  ncu on round 1's kernels (profiles/r01_final_trace_persist_summary.txt, raw page): the per-thread
// local-memory stack made up HALF of the L1 sectors of k_trace_persist (2.7 G loads + 2.9 G stores against 5.8 G sectors
// of node/triangle fetches) and 93 GB of write-through traffic into L2 — lanes of a warp sit at different depths, so one
// STL/LDL touches up to 32 lines — on a kernel whose L1 data pipe runs at 81 %.  The first NS entries of every lane
// therefore live in SHARED memory, laid out [entry][thread]: a lane always hits its own bank whatever its depth, one
// wavefront per warp instruction, nothing leaves the SM.  Deeper entries (rare: ordered traversal keeps the stack short)
// spill to a small per-thread local array.  NS = 0 is the plain local stack (kernel forms that were not converted).
template <int NS>
struct TStack {
    int* sm;   // &s_stack[0][threadIdx.x], stride PT_BLOCK
    int* loc;  // local overflow, PT_STACK - NS entries (PT_STACK_WIDE - NS for the wide walk)
};
template <int NS>
PT_DEV void st_push(const TStack<NS>& S, int& sp, int v) {
    if (NS == 0) S.loc[sp] = v;
    else if (sp < NS) S.sm[sp * PT_BLOCK] = v;
    else S.loc[sp - NS] = v;
    ++sp;
}
template <int NS>
PT_DEV int st_pop(const TStack<NS>& S, int& sp) {
    --sp;
    if (NS == 0) return S.loc[sp];
    return sp < NS ? S.sm[sp * PT_BLOCK] : S.loc[sp - NS];
}

struct Trav {
    float3 inv, oi;  // slab operands: t = box * inv + oi
    float best;      // closest accepted t so far (tmax before any hit)
    Hit h;
    int cur;         // >= 0 inner node, < 0 leaf ~prim, PT_SENTINEL = finished
    int sp;
#ifdef PT_OPT_POSTPONE
    int post;        // a leaf (~prim < 0) met while walking and not yet tested, 0 = none (Aila & Laine's postponed leaf)
#endif
};
#ifdef PT_OPT_POSTPONE
#define PT_TRAV_DONE(T) ((T).cur == PT_SENTINEL && (T).post == 0)
#define PT_TRAV_LEAFWORK(T) ((T).cur < 0 || (T).post != 0)
#else
#define PT_TRAV_DONE(T) ((T).cur == PT_SENTINEL)
#define PT_TRAV_LEAFWORK(T) ((T).cur < 0)
#endif

// A segment begins in two halves: trav_prep tests what is not in the tree (inline spheres from the constant bank, the
// few "global" primitives), trav_start sets up the walk of the tree with that result as the first bound.  The
// persistent kernels do both at once (trav_begin); dual.cu prepares a ray where it is made (shading / regeneration
// phases, many lanes) and starts it when its lane gets to it.
template <bool COUNT>
PT_DEV void trav_prep(const SceneView& sv, float3 o, float3 d, float tmin, float tmax, float& best, Hit& h,
                      TraceCounters& tc) {
    h.t = -1.0f; h.prim = -1; h.u = 0.0f; h.v = 0.0f;
    best = tmax;
    for (int g = 0; g < sv.n_inl; ++g) {  // constant-bank operands, indexed
        if (COUNT) tc.prims++;
        float t;
        const int p = sv.inl_id[g];
        if (sphere_hit_ref(o, d, sv.inl_cr[g], sv.inl_r2[g], sv.inl_transparent[g] != 0, tmin, &t) && t >= tmin &&
            (t < best || (t == best && p < h.prim))) {
            best = t;
            h.t = t; h.prim = p;
        }
    }
    test_inline_tris<COUNT>(sv, o, d, tmin, h, best, tc);
    for (int g = 0; g < sv.n_global; ++g) test_prim<COUNT>(sv, __ldg(&sv.global_prims[g]), o, d, tmin, h, best, tc);
}

template <bool COUNT, int NS>
PT_DEV void trav_start(const SceneView& sv, float3 o, float3 d, float tmin, Trav& T, const TStack<NS>& stack, TraceCounters& tc) {
    T.cur = sv.root;
#ifdef PT_OPT_POSTPONE
    T.post = 0;
#endif
    if (T.cur == PT_SENTINEL) return;  // tree-less scene (<= 8 primitives): the inline / global tests were everything
    T.inv = f3(1.0f / (fabsf(d.x) < 1e-18f ? copysignf(1e-18f, d.x) : d.x),
               1.0f / (fabsf(d.y) < 1e-18f ? copysignf(1e-18f, d.y) : d.y),
               1.0f / (fabsf(d.z) < 1e-18f ? copysignf(1e-18f, d.z) : d.z));
    T.oi = f3(-o.x * T.inv.x, -o.y * T.inv.y, -o.z * T.inv.z);
#ifdef PT_OPT_ROOT_BOX
    if (T.cur >= 0) {  // rays that miss the box of the whole tree (sky and ground rays of the mesh scenes: most of them) never enter the
        // node phase: one slab test here instead of a vote + a two-box node step + a pop
        const float x0 = fmaf(sv.root_lo[0], T.inv.x, T.oi.x), x1 = fmaf(sv.root_hi[0], T.inv.x, T.oi.x);
        const float y0 = fmaf(sv.root_lo[1], T.inv.y, T.oi.y), y1 = fmaf(sv.root_hi[1], T.inv.y, T.oi.y);
        const float z0 = fmaf(sv.root_lo[2], T.inv.z, T.oi.z), z1 = fmaf(sv.root_hi[2], T.inv.z, T.oi.z);
        const float t0 = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
        const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
        if (!(t0 <= t1)) { T.cur = PT_SENTINEL; T.sp = 1; return; }
    }
#endif
    T.sp = 0;
    st_push(stack, T.sp, PT_SENTINEL);
    if (T.cur < 0) {  // degenerate tree: the root is a leaf
        test_prim<COUNT>(sv, ~T.cur, o, d, tmin, T.h, T.best, tc);
        T.cur = PT_SENTINEL;
    }
}

template <bool COUNT, int NS>
PT_DEV void trav_begin(const SceneView& sv, float3 o, float3 d, float tmin, float tmax, Trav& T, const TStack<NS>& stack,
                       TraceCounters& tc) {
    trav_prep<COUNT>(sv, o, d, tmin, tmax, T.best, T.h, tc);
    trav_start<COUNT>(sv, o, d, tmin, T, stack, tc);
}

// 64-byte BVH2 node as two 256-bit loads (LDG.E.ENL2.256, new with sm_100): the L1 spends one tag
// lookup per lane and instruction on these divergent fetches, so two wide loads cost half of four LDG.128
// (profiles/r01_trace_persist_v2_summary.txt: l1tex throughput 96 % with four).
struct __align__(32) Node8 {
    float v[8];
};
PT_DEV Node8 ldg256(const void* p) {
    Node8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
        : "l"(p));
    return r;
}

#define PT_IS_INNER(c) ((unsigned)(c) < (unsigned)PT_SENTINEL)

// What a lane does with the two slab-test results of an inner node: both children hit -> enter the nearer one, push the
// farther; one hit -> enter it; none -> pop.  The three cases are three divergent paths of a warp whose lanes usually take
// all of them.  With the local stack (render kernels) they are replaced by straight-line code: the top of the stack is
// read speculatively, the farther child is stored into the free slot above it, and selects pick the outcome.  The
// batch kernel does the same on its SHARED stack (with a local stack the unconditional accesses thrash L1 there:
// 1672 -> 297 Mrays/s).
template <int NS>
PT_DEV void st_descend(const TStack<NS>& stack, Trav& T, bool ha, bool hb, float t0a, float t0b, int ca, int cb) {
#ifndef PT_OPT_BRANCHY_DESCEND
    if (NS == 0) {  // render kernels: Yoimiya 4949 -> 5230 Mpaths/s (+5.7 %), Zhongli 4K +5.8 %, 10_final +0.8 % (profiles/r02_ab_branchless.txt)
        const bool b_first = hb && (!ha || t0b < t0a);   // both hit: a first iff t0a <= t0b, as the branchy form
        const int near = b_first ? cb : ca, far = b_first ? ca : cb;
        const int top = stack.loc[T.sp - 1];             // sp >= 1: the sentinel sits at the bottom
        stack.loc[T.sp] = far;                           // the slot above the top is free (trees are <= 60 levels deep)
        T.cur = (ha || hb) ? near : top;
        T.sp += (int)(ha && hb) - (int)!(ha || hb);
        return;
    }
#endif
#ifndef PT_OPT_BRANCHY_DESCEND
    if (NS > 0) {  // the same with the shared stack (batch kernel: 1669 -> 1753 Mrays/s, +5.0 %): both memory forms predicated
        const bool b_first = hb && (!ha || t0b < t0a);
        const int near = b_first ? cb : ca, far = b_first ? ca : cb;
        const int sp = T.sp;
        const int top = sp - 1 < NS ? stack.sm[(sp - 1) * PT_BLOCK] : stack.loc[sp - 1 - NS];
        if (sp < NS) stack.sm[sp * PT_BLOCK] = far;
        else stack.loc[sp - NS] = far;
        T.cur = (ha || hb) ? near : top;
        T.sp = sp + (int)(ha && hb) - (int)!(ha || hb);
        return;
    }
#endif
    if (ha && hb) {
        const bool a_first = t0a <= t0b;
        st_push(stack, T.sp, a_first ? cb : ca);
        T.cur = a_first ? ca : cb;
    } else if (ha || hb) {
        T.cur = ha ? ca : cb;
    } else {
        T.cur = st_pop(stack, T.sp);
    }
}

// A lane that has just descended onto a leaf waits for the warp's next service before the primitive is tested
// (persist.cu): its operands are asked into L1 meanwhile (PT_OPT_PREFETCH_LEAF).
PT_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
PT_DEV void prefetch_prim(const SceneView& sv, int p) {
    if (p < sv.n_sph) {
        prefetch_l1(sv.sph_cr + p);
        prefetch_l1(sv.sph_aux + p);
    } else {
        const float4* g = sv.tri_geo + 3 * (size_t)(p - sv.n_sph);
        prefetch_l1(g);
        prefetch_l1(g + 2);  // a 48-byte record may straddle two 128-byte lines
    }
}

// One inner-node step of the ordered traversal: tests both child boxes of node T.cur, descends into the
// nearer hit child (pushing the farther one) or pops.  Afterwards T.cur is an inner node, a leaf (< 0) or
// PT_SENTINEL.
template <bool COUNT, int NS>
PT_DEV void node_step(const SceneView& sv, Trav& T, const TStack<NS>& stack, TraceCounters& tc) {
    if (COUNT) tc.nodes++;
    const float4* n = sv.nodes + 4 * (size_t)T.cur;
    const Node8 A = ldg256(n), B = ldg256(n + 2);
    // child 0: min A0 A1 A2 max A3 A4 A5; child 1: min A6 A7 B0 max B1 B2 B3; refs B4 B5
    const float3 inv = T.inv, oi = T.oi;
    float x0 = fmaf(A.v[0], inv.x, oi.x), x1 = fmaf(A.v[3], inv.x, oi.x);
    float y0 = fmaf(A.v[1], inv.y, oi.y), y1 = fmaf(A.v[4], inv.y, oi.y);
    float z0 = fmaf(A.v[2], inv.z, oi.z), z1 = fmaf(A.v[5], inv.z, oi.z);
    const float t0a = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float t1a = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
    x0 = fmaf(A.v[6], inv.x, oi.x); x1 = fmaf(B.v[1], inv.x, oi.x);
    y0 = fmaf(A.v[7], inv.y, oi.y); y1 = fmaf(B.v[2], inv.y, oi.y);
    z0 = fmaf(B.v[0], inv.z, oi.z); z1 = fmaf(B.v[3], inv.z, oi.z);
    const float t0b = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float t1b = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
    const bool ha = t0a <= t1a, hb = t0b <= t1b;
    const int ca = __float_as_int(B.v[4]), cb = __float_as_int(B.v[5]);
    st_descend(stack, T, ha, hb, t0a, t0b, ca, cb);
#ifdef PT_OPT_PREFETCH_LEAF
    if (T.cur < 0) prefetch_prim(sv, ~T.cur);
#endif
#ifdef PT_OPT_POSTPONE
    // the first leaf a lane lands on is set aside and the walk goes on (the lane stays in the node phase); the second
    // one makes it wait for the warp's next service, where both are tested
    if (T.cur < 0 && T.post == 0) { T.post = T.cur; T.cur = st_pop(stack, T.sp); }
#endif
}

// The same step on the 32-byte quantised nodes (one 256-bit load): the child boxes are 12 x u16 in the frame of the
// root box, and the ray's slab operands are pre-multiplied by that frame (trav_frame_q), so the decode is just the
// twelve integer -> float conversions (I2F.U16 with half-register selectors; an ALU-only decode via the 2^23 + q
// float trick measured the same, so the XU pipe is not what limits this kernel).
PT_DEV void trav_frame_q(const SceneView& sv, float3 o, Trav& T) {  // after trav_begin: t = q * (scale * inv) + (lo - o) * inv
    T.oi = f3((sv.qlo[0] - o.x) * T.inv.x, (sv.qlo[1] - o.y) * T.inv.y, (sv.qlo[2] - o.z) * T.inv.z);
    T.inv = f3(sv.qscale[0] * T.inv.x, sv.qscale[1] * T.inv.y, sv.qscale[2] * T.inv.z);
}
struct __align__(32) QNode8 {
    unsigned w[8];
};
PT_DEV QNode8 ldg256u(const void* p) {
    QNode8 r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
        : "l"(p));
    return r;
}
PT_DEV float qlo16(unsigned w) { return (float)(unsigned short)w; }
PT_DEV float qhi16(unsigned w) { return (float)(unsigned short)(w >> 16); }

template <bool COUNT, int NS>
PT_DEV void node_step_q(const SceneView& sv, Trav& T, const TStack<NS>& stack, TraceCounters& tc) {
    if (COUNT) tc.nodes++;
    const QNode8 N = ldg256u(sv.qnodes + 2 * (size_t)T.cur);
    // w0 = c0.min x|y, w1 = c0.min z | c0.max x, w2 = c0.max y|z, w3 = c1.min x|y, w4 = c1.min z | c1.max x, w5 = c1.max y|z
    const float3 inv = T.inv, oi = T.oi;
    float x0 = fmaf(qlo16(N.w[0]), inv.x, oi.x), x1 = fmaf(qhi16(N.w[1]), inv.x, oi.x);
    float y0 = fmaf(qhi16(N.w[0]), inv.y, oi.y), y1 = fmaf(qlo16(N.w[2]), inv.y, oi.y);
    float z0 = fmaf(qlo16(N.w[1]), inv.z, oi.z), z1 = fmaf(qhi16(N.w[2]), inv.z, oi.z);
    const float t0a = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float t1a = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
    x0 = fmaf(qlo16(N.w[3]), inv.x, oi.x); x1 = fmaf(qhi16(N.w[4]), inv.x, oi.x);
    y0 = fmaf(qhi16(N.w[3]), inv.y, oi.y); y1 = fmaf(qlo16(N.w[5]), inv.y, oi.y);
    z0 = fmaf(qlo16(N.w[4]), inv.z, oi.z); z1 = fmaf(qhi16(N.w[5]), inv.z, oi.z);
    const float t0b = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float t1b = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
    const bool ha = t0a <= t1a, hb = t0b <= t1b;
    const int ca = (int)N.w[6], cb = (int)N.w[7];
    st_descend(stack, T, ha, hb, t0a, t0b, ca, cb);
#ifdef PT_OPT_PREFETCH_LEAF
    if (T.cur < 0) prefetch_prim(sv, ~T.cur);
#endif
#ifdef PT_OPT_POSTPONE
    // the first leaf a lane lands on is set aside and the walk goes on (the lane stays in the node phase); the second
    // one makes it wait for the warp's next service, where both are tested
    if (T.cur < 0 && T.post == 0) { T.post = T.cur; T.cur = st_pop(stack, T.sp); }
#endif
}

// EXPERIMENTAL (opt-in, see bvh4.h; bit-identical hit records on the B200, 0.57x the steps, 1.32x on a small batch).
// One step over a 4-wide node (bvh4.h: minx[4] miny[4] | minz[4] maxx[4] | maxy[4] maxz[4] | ref[4] pad[4], three 256-bit
// loads + one 128-bit): four slab tests, the nearest hit child is entered, the other hit children are pushed in slot
// order (CPU prototype: full sorting would save only 2 % of the steps).  Pushes up to three entries per step: the
// kernel that uses it carries a 128-entry stack.
#define PT_WIDE_EMPTY 0x7fffffff
#define PT_STACK_WIDE 128
template <bool COUNT, int NS>
PT_DEV void node_step4(const SceneView& sv, Trav& T, const TStack<NS>& stack, TraceCounters& tc) {
    if (COUNT) tc.nodes++;
    const float4* n = sv.wnodes + 8 * (size_t)T.cur;
    const Node8 A = ldg256(n), B = ldg256(n + 2), C = ldg256(n + 4);
    const int4 R = __ldg((const int4*)(n + 6));
    const float3 inv = T.inv, oi = T.oi;
    const int ref[4] = {R.x, R.y, R.z, R.w};
    float t0[4];
    bool h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x0 = fmaf(A.v[k], inv.x, oi.x), x1 = fmaf(B.v[4 + k], inv.x, oi.x);
        const float y0 = fmaf(A.v[4 + k], inv.y, oi.y), y1 = fmaf(C.v[k], inv.y, oi.y);
        const float z0 = fmaf(B.v[k], inv.z, oi.z), z1 = fmaf(C.v[4 + k], inv.z, oi.z);
        t0[k] = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
        const float t1 = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), T.best));
        h[k] = t0[k] <= t1 && ref[k] != PT_WIDE_EMPTY;
    }
    int near = -1;
    float tn = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (h[k] && (near < 0 || t0[k] < tn)) { near = k; tn = t0[k]; }
    int next = PT_SENTINEL;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (h[k]) {
            if (k == near) next = ref[k];
            else st_push(stack, T.sp, ref[k]);
        }
    }
    if (near >= 0) T.cur = next;
    else T.cur = st_pop(stack, T.sp);
}

// Leaf step: tests primitive ~T.cur and pops.
template <bool COUNT, int NS>
PT_DEV void leaf_step(const SceneView& sv, float3 o, float3 d, float tmin, Trav& T, const TStack<NS>& stack, TraceCounters& tc) {
#ifdef PT_OPT_POSTPONE
    if (T.post != 0) {
        test_prim<COUNT>(sv, ~T.post, o, d, tmin, T.h, T.best, tc);
        T.post = 0;
    }
    if (T.cur < 0) {
        test_prim<COUNT>(sv, ~T.cur, o, d, tmin, T.h, T.best, tc);
        T.cur = st_pop(stack, T.sp);
        if (T.cur < 0) { T.post = T.cur; T.cur = st_pop(stack, T.sp); }
    }
#else
    test_prim<COUNT>(sv, ~T.cur, o, d, tmin, T.h, T.best, tc);
    T.cur = st_pop(stack, T.sp);
#endif
}
