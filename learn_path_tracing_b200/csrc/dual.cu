// dual.cu — PT_MODE_DUAL: the persistent ballot-scheduled kernel of persist.cu with TWO paths per lane.
//
// persist.cu's lanes wait with a finished traversal until enough of them have piled up to be shaded, and a shaded
// group splits into hits and misses (ncu: node steps at 16-24 of 32 lanes, shading at ~14).  queue.cu decoupled the two
// through block-wide queues and paid more for the atomics than it gained.  Here the decoupling is LANE-PRIVATE: every
// lane owns two path records in shared memory (its own 2 x 64 bytes: no indices, no atomics, no bank conflicts) and
// walks with one of them — only the ray and the traversal state are in registers; throughput, pixel and sample stay in
// the record.  A finished traversal just writes its hit into the record and the lane flips to its other record if a
// ray waits there; hits, misses and empty records are then served by three separately voted warp phases, each only
// once enough lanes want exactly it — hits and misses no longer share an instruction stream, and nobody idles:
//
//   finish    traversal done: hit record -> S.h, record becomes HIT or MISS
//   miss      MISS records: sky / environment * throughput -> RED.v4, record becomes EMPTY
//   hit       HIT records: scatter; the continued ray is written back (RAW), or EMPTY at the depth limit
//   regen     EMPTY records get the next camera path of the warp's work unit (fused ray generation), RAW
//   prep      RAW records: inline spheres + global primitives are tested (first half of the segment), READY —
//             or, in a tree-less scene, HIT / MISS at once: such scenes never walk, they only cycle these phases
//   start     lanes that are not walking flip to a READY record and walk the tree for it
//
// Same RNG keys (pixel, sample, bounce) as every other mode: the same set of paths.
#include <math.h>
#include <string.h>

#include "wf_common.cuh"

#define D_BLOCK 256
#define D_TILE_W 8
#define D_TILE_H 4
#define D_UNIT_SAMPLES 16

#define PK_EMPTY 0
#define PK_RAW 1    // holds a ray (camera or scattered) whose segment has not been prepared yet
#define PK_READY 2  // prepared (inline spheres / global primitives tested, result in S.h): waits to be walked
#define PK_HIT 3
#define PK_MISS 4
#define PK_WALK 5

struct DShared {  // two lane-private path records, one float4 column per field: conflict-free 128-bit accesses
    float4 a[2 * D_BLOCK];  // o.xyz | bits(pixel)
    float4 b[2 * D_BLOCK];  // d.xyz | bits(sample | bounce << 24)
    float4 c[2 * D_BLOCK];  // throughput.rgb | -
    float4 h[2 * D_BLOCK];  // t, bits(prim), u, v
};

template <bool LEGACY, bool COUNT, int MINB>
__global__ void __launch_bounds__(D_BLOCK, MINB)
k_paths_dual(const SceneView sv, const RenderConsts rc, unsigned long long* __restrict__ counters,
             float4* __restrict__ accum, float4* __restrict__ accum_sq, int shade_min, int serve_min) {
    __shared__ DShared S;
    int lstack[PT_STACK];
    const TStack<0> stack = {nullptr, lstack};  // experimental kernel form: plain local-memory stack
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned tiles_x = ((unsigned)rc.W + D_TILE_W - 1) / D_TILE_W, tiles_y = ((unsigned)rc.H + D_TILE_H - 1) / D_TILE_H;
    const unsigned spp = rc.sample_end - rc.spp_offset;
    const unsigned n_chunks = (spp + D_UNIT_SAMPLES - 1) / D_UNIT_SAMPLES;
    const unsigned long long n_units = (unsigned long long)tiles_x * tiles_y * n_chunks;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);  // the ray being walked (copy of record `ci`)
    Trav T;
    T.cur = PT_SENTINEL; T.sp = 1;
    unsigned ci = tid;             // index of the current record (tid or tid + D_BLOCK)
    int sc = PK_EMPTY, so = PK_EMPTY;  // state of the current / the other record; only the current one can be WALK
    unsigned unit_x0 = 0, unit_y0 = 0, unit_s0 = 0, unit_next = 0, unit_size = 0;  // warp-uniform work unit
    bool exhausted = false;
    int walk_min = 0;
    unsigned nseg = 0;
    TraceCounters tc;
    tc.nodes = 0; tc.prims = 0;

    for (;;) {
        const bool inner = PT_IS_INNER(T.cur);  // T.cur == PT_SENTINEL unless sc == PK_WALK
        const int n_inner = __popc(__ballot_sync(0xffffffffu, inner));
        if (n_inner > walk_min) {
            if (inner) {
                node_step<COUNT>(sv, T, stack, tc);
                if (n_inner > walk_min + 4 && PT_IS_INNER(T.cur)) node_step<COUNT>(sv, T, stack, tc);
            }
            continue;
        }
        // ---- service -------------------------------------------------------------------------------------------
        if (T.cur < 0) leaf_step<COUNT>(sv, o, d, rc.tmin, T, stack, tc);
        __syncwarp();
        if (sc == PK_WALK && T.cur == PT_SENTINEL) {  // finish: the hit goes into the record
            S.h[ci] = make_float4(T.best, __int_as_float(T.h.prim), T.h.u, T.h.v);
            sc = T.h.prim >= 0 ? PK_HIT : PK_MISS;
        }
        // phases fire once shade_min lanes want them; with hardly anybody walking, whatever is there
        const int n_walk = __popc(__ballot_sync(0xffffffffu, sc == PK_WALK));
        const int thr = n_walk >= 8 ? shade_min : 1;
        {
            const bool want = sc == PK_MISS || so == PK_MISS;
            if (__popc(__ballot_sync(0xffffffffu, want)) >= thr) {
                if (want) {  // miss: sky / environment radiance * throughput into the accumulator, path ends
                    const bool cur = sc == PK_MISS;
                    const unsigned k = cur ? ci : ci ^ D_BLOCK;
                    const float4 qa = S.a[k], qb = S.b[k], qc = S.c[k];
                    const float3 c = (LEGACY ? environment_color(sv, f3(qb)) : sky_color(f3(qb))) * f3(qc);
                    const unsigned pixel = __float_as_uint(qa.w);
                    if (isfinite(c.x) && isfinite(c.y) && isfinite(c.z)) {
                        atomicAdd(&accum[pixel], make_float4(c.x, c.y, c.z, 1.0f));
                        if (rc.accum_sq) atomicAdd(&accum_sq[pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                    }
                    if (cur) sc = PK_EMPTY; else so = PK_EMPTY;
                }
                __syncwarp();
            }
        }
        {
            const bool want = sc == PK_HIT || so == PK_HIT;
            if (__popc(__ballot_sync(0xffffffffu, want)) >= thr) {
                if (want) {
                    const bool cur = sc == PK_HIT;
                    const unsigned k = cur ? ci : ci ^ D_BLOCK;
                    const float4 qa = S.a[k], qb = S.b[k], qc = S.c[k], qh = S.h[k];
                    PathState q;
                    q.o = f3(qa); q.d = f3(qb); q.l = f3(qc);
                    q.pixel = __float_as_uint(qa.w);
                    const uint32_t sb = __float_as_uint(qb.w);
                    q.sample = sb & 0xFFFFFFu;
                    q.bounce = sb >> 24;
                    Hit hh;
                    hh.t = qh.x; hh.prim = __float_as_int(qh.y); hh.u = qh.z; hh.v = qh.w;
                    int ns = PK_EMPTY;
                    if (!LEGACY && rc.shading_model == PT_SHADE_V2_NORMALS) {  // stages 4-5: normal as colour, no bounce
                        const float3 c = normal_color(sv, q, hh);
                        atomicAdd(&accum[q.pixel], make_float4(c.x, c.y, c.z, 1.0f));
                        if (rc.accum_sq) atomicAdd(&accum_sq[q.pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                    } else {
                        if (LEGACY) scatter_legacy(sv, q, hh, rc.absorptivity, rc.seed, sv.lut);
                        else scatter_v2(sv, q, hh, rc.shading_model, rc.seed);
                        q.bounce += 1u;
                        if (q.bounce < (uint32_t)rc.max_depth) {  // over propagate_limit: contributes nothing
                            S.a[k] = make_float4(q.o.x, q.o.y, q.o.z, qa.w);
                            S.b[k] = make_float4(q.d.x, q.d.y, q.d.z, __uint_as_float(q.sample | (q.bounce << 24)));
                            S.c[k] = make_float4(q.l.x, q.l.y, q.l.z, 0.0f);
                            ns = PK_RAW;
                        }
                    }
                    if (cur) sc = ns; else so = ns;
                }
                __syncwarp();
            }
        }
        if (!exhausted && __popc(__ballot_sync(0xffffffffu, sc == PK_EMPTY || so == PK_EMPTY)) >= thr) {
            // regen: empty records take the next camera paths of the warp's work unit (Camera.get_rays fused);
            // one record per lane and pass
            for (;;) {
                const bool empty = sc == PK_EMPTY || so == PK_EMPTY;
                const unsigned want = __ballot_sync(0xffffffffu, empty);
                if (want == 0u) break;
                if (unit_next >= unit_size) {
                    unsigned long long u = 0ull;
                    if (lane == 0u) u = atomicAdd(&counters[CNT_NEXT_PATH], 1ull);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= n_units) {
                        exhausted = true;
                        break;
                    }
                    const unsigned tile = (unsigned)(u / n_chunks), chunk = (unsigned)(u - (unsigned long long)tile * n_chunks);
                    const unsigned ty = tile / tiles_x;
                    unit_x0 = (tile - ty * tiles_x) * D_TILE_W;
                    unit_y0 = ty * D_TILE_H;
                    unit_s0 = chunk * D_UNIT_SAMPLES;
                    unit_size = min((unsigned)D_UNIT_SAMPLES, spp - unit_s0) * 32u;
                    unit_next = 0u;
                }
                const unsigned take = min((unsigned)__popc(want), unit_size - unit_next);
                if (empty) {
                    const unsigned r = __popc(want & lt);
                    if (r < take) {
                        const unsigned q = unit_next + r;
                        const unsigned px = unit_x0 + (q & 7u), py = unit_y0 + ((q >> 3) & 3u);
                        if (px < (unsigned)rc.W && py < (unsigned)rc.H) {  // image sizes need not be tile multiples
                            const unsigned pixel = py * (unsigned)rc.W + px, sample = rc.spp_offset + unit_s0 + (q >> 5);
                            float3 co, cd;
                            camera_ray(rc.cam, (int)px, (int)py, rng4(pixel, sample, 0u, rc.seed), &co, &cd);
                            const bool cur = sc == PK_EMPTY;
                            const unsigned k = cur ? ci : ci ^ D_BLOCK;
                            S.a[k] = make_float4(co.x, co.y, co.z, __uint_as_float(pixel));
                            S.b[k] = make_float4(cd.x, cd.y, cd.z, __uint_as_float(sample));  // bounce 0
                            S.c[k] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                            if (cur) sc = PK_RAW; else so = PK_RAW;
                        }
                    }
                }
                unit_next += take;
            }
            __syncwarp();
        }
        // prep: every ray made by the two phases above gets the first half of its segment — the inline spheres and the
        // global primitives — HERE, with all the lanes that made one, not later when its lane gets to walk it (ncu on
        // the first version: trav_begin at 7 of 32 lanes, a quarter of all instructions).  In a tree-less scene that is
        // the whole segment: the record becomes HIT or MISS at once.
        for (;;) {
            const bool want = sc == PK_RAW || so == PK_RAW;
            if (__ballot_sync(0xffffffffu, want) == 0u) break;
            if (want) {
                const bool cur = sc == PK_RAW;
                const unsigned k = cur ? ci : ci ^ D_BLOCK;
                const float4 qa = S.a[k], qb = S.b[k];
                float best;
                Hit hh;
                trav_prep<COUNT>(sv, f3(qa), f3(qb), rc.tmin, INFINITY, best, hh, tc);
                ++nseg;
                S.h[k] = make_float4(best, __int_as_float(hh.prim), hh.u, hh.v);
                const int ns = sv.root != PT_SENTINEL ? PK_READY : (hh.prim >= 0 ? PK_HIT : PK_MISS);
                if (cur) sc = ns; else so = ns;
            }
            __syncwarp();
        }
        // start: a lane that is not walking flips to a READY record and walks the tree for it
        if (sc != PK_WALK && (sc == PK_READY || so == PK_READY)) {
            if (sc != PK_READY) {
                ci ^= D_BLOCK;
                const int t = sc; sc = so; so = t;
            }
            const float4 qa = S.a[ci], qb = S.b[ci], qh = S.h[ci];
            o = f3(qa); d = f3(qb);
            T.best = qh.x;
            T.h.t = qh.x; T.h.prim = __float_as_int(qh.y); T.h.u = qh.z; T.h.v = qh.w;
            trav_start<COUNT>(sv, o, d, rc.tmin, T, stack, tc);
            sc = PK_WALK;
        }
        __syncwarp();
        if (exhausted && __ballot_sync(0xffffffffu, (sc | so) != PK_EMPTY) == 0u) break;  // nothing left anywhere
        walk_min = max(0, __popc(__ballot_sync(0xffffffffu, T.cur != PT_SENTINEL)) - serve_min);
    }
    nseg += __shfl_xor_sync(0xffffffffu, nseg, 16);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, 8);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, 4);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, 2);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, 1);
    if (lane == 0 && nseg) atomicAdd(&counters[CNT_SEGMENTS], (unsigned long long)nseg);
    if (COUNT) {
        atomicAdd(&counters[CNT_NODES], (unsigned long long)tc.nodes);
        atomicAdd(&counters[CNT_PRIMS], (unsigned long long)tc.prims);
    }
}

template <bool LEGACY, bool COUNT, int MINB>
static int launch_dual(PtContext* ctx, const PtScene* s, const RenderConsts& rc, float4* accum, float4* accum_sq, int shade_min,
                       int serve_min) {
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_paths_dual<LEGACY, COUNT, MINB>, D_BLOCK, 0));
    if (per_sm < 1) per_sm = 1;
    int blocks = per_sm * ctx->sm_count;
    const unsigned long long need = (rc.total_paths + 2 * D_BLOCK - 1) / (2 * D_BLOCK);
    if ((unsigned long long)blocks > need) blocks = (int)need;
    if (blocks < 1) return PT_OK;
    k_paths_dual<LEGACY, COUNT, MINB><<<blocks, D_BLOCK, 0, ctx->stream>>>(s->view, rc, ctx->counters, accum, accum_sq, shade_min, serve_min);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

// blocks_per_sm: 3 (85 registers, no spills) or 4 (64 registers)
int pt_render_dual(PtContext* ctx, const PtScene* s, const RenderConsts& rc, bool legacy, bool count, float4* accum,
                   float4* accum_sq, int shade_min, int serve_min, int blocks_per_sm) {
#define DUAL_GO(LG, CN)                                                                                         \
    return blocks_per_sm == 3 ? launch_dual<LG, CN, 3>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)       \
                              : launch_dual<LG, CN, 4>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
    if (legacy) {
        if (count) DUAL_GO(true, true);
        DUAL_GO(true, false);
    }
    if (count) DUAL_GO(false, true);
    DUAL_GO(false, false);
#undef DUAL_GO
}
