// queue.cu — PT_MODE_QUEUE: the persistent ballot-scheduled kernel of persist.cu with the shading decoupled from the
// traversal through BLOCK-LOCAL queues in shared memory.
//
// ncu on k_paths_persist (profiles/r01_paths_persist_*): node steps run at 16-24 of 32 lanes because finished lanes
// wait in their warp until enough of them have piled up to be shaded, and the shading itself runs at ~14 lanes because
// every group splits into hits and misses.  Here a path does not belong to a lane any more:
//
//   slot      384 path records per block in shared memory (origin|pixel, direction|sample+bounce, throughput, hit)
//   q_ready   slots holding a ray that waits to be traversed
//   q_hit     slots whose traversal ended on a primitive      } lock-free index rings in shared memory,
//   q_miss    slots whose ray left the scene                   } warp-aggregated push / pop (one atomic per warp)
//
//   a lane    takes a slot from q_ready, walks the BVH for it (node / leaf phases voted by ballot as in persist.cu;
//             only the traversal state lives in registers, origin and direction are re-read from the slot for the rare
//             leaf tests), writes the hit record into the slot, pushes the slot to q_hit or q_miss and immediately
//             takes the next ready slot: no lane waits for shading
//   a warp    that sees 32 entries in q_hit (or q_miss) pops them and shades them with ALL lanes — hits and misses
//             never share an instruction stream — writes the scattered ray (or, when the path ended, the next camera
//             ray of the block's work unit: fused ray generation) back into the same slot and pushes it to q_ready
//
// (A first version with index stacks under one spin lock per block ran 3-6x SLOWER than persist.cu: eight warps
// serialised on the lock.  Hence the lock-free rings and per-warp work units.)
//
// MEASURED RESULT (B200, profiles/r01_paths_queue_final_v1_summary.txt): the idea works as intended — node steps run at
// 28 of 32 lanes (persist.cu: 24) and shading at ~21 (persist.cu: ~14) — but the machinery costs more than it saves:
// 61.6 G warp instructions instead of 40.4 G on 10_final (17 % of them shared-memory atomics and their plumbing, the
// path records travelling through shared memory, trav_begin at the 8-24 lanes a refill brings), so the renders are
// 1.3-1.8x SLOWER (10_final 3.8 vs 5.7 Gpaths/s, Yoimiya 3.3 vs 4.2, 8_refract 9.2 vs 17.0).  The mode is kept as an
// experiment (same paths as every other mode, tested); PT_MODE_PERSIST stays the default.
// The RNG is keyed on (pixel, sample, bounce): same set of paths as every other mode.
#include <math.h>
#include <string.h>

#include "wf_common.cuh"

#define Q_BLOCK 256
#define Q_SLOTS 384
#define Q_TILE_W 8
#define Q_TILE_H 4
#define Q_UNIT_SAMPLES 16

#define Q_RING 512  // ring capacity (power of two, > Q_SLOTS: a slot sits in at most one ring)
#define Q_NONE 0xFFFFu

// Lock-free multi-producer / multi-consumer ring of slot indices.  Producers reserve positions with one atomicAdd per
// warp on `tail`, write the indices, then publish how many entries are complete through `count`; consumers take
// entries off `count` first (so they never reserve more than exists), then positions off `head`, and wait the few
// cycles a producer may still need to fill a reserved position (Q_NONE marks an unfilled one).
struct QRing {
    unsigned short e[Q_RING];
    int head, tail, count;
};

struct QShared {
    float4 a[Q_SLOTS];  // o.xyz | bits(pixel)
    float4 b[Q_SLOTS];  // d.xyz | bits(sample | bounce << 24)
    float4 c[Q_SLOTS];  // throughput.rgb | -
    float4 h[Q_SLOTS];  // t, bits(prim), u, v
    QRing ready, hit, miss, freeq;  // freeq: slots whose path ended when the shading warp had no camera path left to give
    int active;  // live paths of the block
};

#define QV(x) (*(volatile int*)&(x))
// warp-uniform snapshot of a shared counter that other warps change (lanes need not be converged at the read)
#define QUNI(x) __shfl_sync(0xffffffffu, QV(x), 0)

PT_DEV void q_push(QRing& R, bool pred, int slot, unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const int k = __popc(m);
    int t = 0;
    if (lane == 0u) t = atomicAdd(&R.tail, k);
    t = __shfl_sync(0xffffffffu, t, 0);
    __threadfence_block();  // the slot's record (written by this lane before) becomes visible before its index does
    if (pred) *(volatile unsigned short*)&R.e[(t + __popc(m & ((1u << lane) - 1u))) & (Q_RING - 1)] = (unsigned short)slot;
    __syncwarp();
    __threadfence_block();
    if (lane == 0u) atomicAdd(&R.count, k);
}

// Lanes with `want` receive a slot index or -1.
PT_DEV int q_pop(QRing& R, bool want, unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return -1;
    const int k = __popc(m);
    int take = 0, h = 0;
    if (lane == 0u) {
        if (QV(R.count) > 0) {  // cheap look before the atomics
            const int old = atomicSub(&R.count, k);
            take = max(0, min(k, old));
            if (take < k) atomicAdd(&R.count, k - take);
            if (take > 0) h = atomicAdd(&R.head, take);
        }
    }
    take = __shfl_sync(0xffffffffu, take, 0);
    h = __shfl_sync(0xffffffffu, h, 0);
    int slot = -1;
    const int r = __popc(m & ((1u << lane) - 1u));
    if (want && r < take) {
        volatile unsigned short* e = &R.e[(h + r) & (Q_RING - 1)];
        unsigned v;
        while ((v = *e) == Q_NONE) {}
        *e = (unsigned short)Q_NONE;
        slot = (int)v;
    }
    __syncwarp();
    __threadfence_block();
    return slot;
}

// Per-warp work unit (8x4 pixel tile x 16 samples from ONE global counter), kept in warp-uniform registers.
struct QUnit {
    unsigned x0, y0, s0, size, next;
    bool exhausted;
};

// Warp-collective: lanes with `need` get the next camera path of the warp's work unit written into their slot
// (Camera.get_rays fused: camera.py:71-93, 15_module.py:438-453).  Returns false for a lane when no paths are left.
PT_DEV bool q_new_paths(QShared& S, QUnit& U, bool need, int slot, const RenderConsts& rc, unsigned long long* counters,
                        unsigned lane, unsigned tiles_x, unsigned n_chunks, unsigned long long n_units, unsigned spp) {
    bool got = false;
    const unsigned lt = (1u << lane) - 1u;
    for (;;) {
        const unsigned want = __ballot_sync(0xffffffffu, need && !got);
        if (want == 0u) break;
        if (U.next >= U.size) {
            if (U.exhausted) break;
            unsigned long long u = 0ull;
            if (lane == 0u) u = atomicAdd(&counters[CNT_NEXT_PATH], 1ull);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= n_units) {
                U.exhausted = true;
                break;
            }
            const unsigned tile = (unsigned)(u / n_chunks), chunk = (unsigned)(u - (unsigned long long)tile * n_chunks);
            const unsigned ty = tile / tiles_x;
            U.x0 = (tile - ty * tiles_x) * Q_TILE_W;
            U.y0 = ty * Q_TILE_H;
            U.s0 = chunk * Q_UNIT_SAMPLES;
            U.size = min((unsigned)Q_UNIT_SAMPLES, spp - U.s0) * 32u;
            U.next = 0u;
        }
        const unsigned take = min((unsigned)__popc(want), U.size - U.next);
        if (need && !got) {
            const unsigned r = __popc(want & lt);
            if (r < take) {
                const unsigned q = U.next + r;
                const unsigned px = U.x0 + (q & 7u), py = U.y0 + ((q >> 3) & 3u);
                if (px < (unsigned)rc.W && py < (unsigned)rc.H) {  // image sizes need not be tile multiples
                    const unsigned pixel = py * (unsigned)rc.W + px, sample = rc.spp_offset + U.s0 + (q >> 5);
                    float3 o, d;
                    camera_ray(rc.cam, (int)px, (int)py, rng4(pixel, sample, 0u, rc.seed), &o, &d);
                    S.a[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pixel));
                    S.b[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(sample));  // bounce 0
                    S.c[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                    got = true;
                }
            }
        }
        U.next += take;
    }
    return got;
}

template <bool LEGACY, bool COUNT>
__global__ void __launch_bounds__(Q_BLOCK, 4)
k_paths_queue(const SceneView sv, const RenderConsts rc, unsigned long long* __restrict__ counters,
              float4* __restrict__ accum, float4* __restrict__ accum_sq, int serve_min) {
    __shared__ QShared S;
    int lstack[PT_STACK];
    const TStack<0> stack = {nullptr, lstack};  // experimental kernel form: plain local-memory stack
    const unsigned lane = threadIdx.x & 31u;
    const unsigned tiles_x = ((unsigned)rc.W + Q_TILE_W - 1) / Q_TILE_W, tiles_y = ((unsigned)rc.H + Q_TILE_H - 1) / Q_TILE_H;
    const unsigned spp = rc.sample_end - rc.spp_offset;
    const unsigned n_chunks = (spp + Q_UNIT_SAMPLES - 1) / Q_UNIT_SAMPLES;
    const unsigned long long n_units = (unsigned long long)tiles_x * tiles_y * n_chunks;
    for (int i = threadIdx.x; i < Q_RING; i += Q_BLOCK) S.ready.e[i] = S.hit.e[i] = S.miss.e[i] = S.freeq.e[i] = (unsigned short)Q_NONE;
    if (threadIdx.x == 0) {
        S.ready.head = S.ready.tail = S.ready.count = 0;
        S.hit.head = S.hit.tail = S.hit.count = 0;
        S.miss.head = S.miss.tail = S.miss.count = 0;
        S.freeq.head = S.freeq.tail = S.freeq.count = 0;
        S.active = 0;
    }
    QUnit U;
    U.x0 = U.y0 = U.s0 = U.size = U.next = 0u;
    U.exhausted = false;
    __syncthreads();
    // ---- every slot starts with a camera path ----------------------------------------------------------------
    for (int base = 0; base < Q_SLOTS; base += Q_BLOCK) {
        const int slot = base + (int)threadIdx.x;
        const bool need = slot < Q_SLOTS;
        const bool got = q_new_paths(S, U, need, slot, rc, counters, lane, tiles_x, n_chunks, n_units, spp);
        const unsigned mg = __ballot_sync(0xffffffffu, got);
        if (lane == 0u && mg) atomicAdd(&S.active, __popc(mg));
        q_push(S.ready, got, slot, lane);
        q_push(S.freeq, need && !got, slot, lane);
    }
    __syncthreads();

    Trav T;
    T.cur = PT_SENTINEL; T.sp = 1;
    int slot = -1;     // the slot this lane traverses for (-1: none)
    int walk_min = 0;  // the node phase runs while more than walk_min lanes stand on an inner node
    int n_empty = 32;  // lanes without a slot after the last service
    int idle = 0;
    unsigned nseg = 0;
    TraceCounters tc;
    tc.nodes = 0; tc.prims = 0;

    for (;;) {
        const bool inner = PT_IS_INNER(T.cur);  // T.cur == PT_SENTINEL whenever slot < 0
        const int n_inner = __popc(__ballot_sync(0xffffffffu, inner));
        if (n_inner > walk_min && !(n_empty > 0 && QUNI(S.ready.count) > 0)) {
            if (inner) {
                node_step<COUNT>(sv, T, stack, tc);
                if (n_inner > walk_min + 4 && PT_IS_INNER(T.cur)) node_step<COUNT>(sv, T, stack, tc);
            }
            continue;
        }
        // ---- service -------------------------------------------------------------------------------------------
        if (T.cur < 0) {  // leaf: origin and direction come back from the slot
            const float4 ra = S.a[slot], rb = S.b[slot];
            leaf_step<COUNT>(sv, f3(ra), f3(rb), rc.tmin, T, stack, tc);
        }
        __syncwarp();
        {  // finished traversals: hit record into the slot, slot onto q_hit / q_miss
            const bool fin = slot >= 0 && T.cur == PT_SENTINEL;
            if (fin) S.h[slot] = make_float4(T.h.prim >= 0 ? T.best : -1.0f, __int_as_float(T.h.prim), T.h.u, T.h.v);
            q_push(S.hit, fin && T.h.prim >= 0, slot, lane);
            q_push(S.miss, fin && T.h.prim < 0, slot, lane);
            if (fin) slot = -1;
        }
        {  // lanes without a slot take the next ready ray and start its traversal
            const bool want = slot < 0;
            const int s2 = q_pop(S.ready, want, lane);
            if (want && s2 >= 0) {
                slot = s2;
                const float4 ra = S.a[slot], rb = S.b[slot];
                trav_begin<COUNT>(sv, f3(ra), f3(rb), rc.tmin, INFINITY, T, stack, tc);
                ++nseg;
            }
        }
        __syncwarp();
        n_empty = __popc(__ballot_sync(0xffffffffu, slot < 0));
        // ---- shading: a full group of hits or of misses; anything at all when this warp has lanes to spare -------
        const int nh = QUNI(S.hit.count), nm = QUNI(S.miss.count);
        int what = 0;  // 1 = hits, 2 = misses
        if (nh >= 32 || nm >= 32) what = nh >= nm ? 1 : 2;
        else if (n_empty > 0 && (nh > 0 || nm > 0) && (n_empty == 32 || nh + nm >= 24 || idle > 2)) what = nh >= nm ? 1 : 2;
        if (what == 1) {
            const int s = q_pop(S.hit, true, lane);
            bool alive = false, regen = false;
            if (s >= 0) {
                const float4 ra = S.a[s], rb = S.b[s], rl = S.c[s], rh = S.h[s];
                PathState p;
                p.o = f3(ra); p.d = f3(rb); p.l = f3(rl);
                p.pixel = __float_as_uint(ra.w);
                const uint32_t sb = __float_as_uint(rb.w);
                p.sample = sb & 0xFFFFFFu;
                p.bounce = sb >> 24;
                Hit hh;
                hh.t = rh.x; hh.prim = __float_as_int(rh.y); hh.u = rh.z; hh.v = rh.w;
                if (!LEGACY && rc.shading_model == PT_SHADE_V2_NORMALS) {  // stages 4-5: normal as colour, no bounce
                    const float3 c = normal_color(sv, p, hh);
                    atomicAdd(&accum[p.pixel], make_float4(c.x, c.y, c.z, 1.0f));
                    if (rc.accum_sq) atomicAdd(&accum_sq[p.pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                } else {
                    if (LEGACY) scatter_legacy(sv, p, hh, rc.absorptivity, rc.seed, sv.lut);
                    else scatter_v2(sv, p, hh, rc.shading_model, rc.seed);
                    p.bounce += 1u;
                    alive = p.bounce < (uint32_t)rc.max_depth;  // over propagate_limit: contributes nothing
                }
                if (alive) {
                    S.a[s] = make_float4(p.o.x, p.o.y, p.o.z, ra.w);
                    S.b[s] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.sample | (p.bounce << 24)));
                    S.c[s] = make_float4(p.l.x, p.l.y, p.l.z, 0.0f);
                } else {
                    regen = true;
                }
            }
            const bool got = q_new_paths(S, U, regen, s, rc, counters, lane, tiles_x, n_chunks, n_units, spp);
            const unsigned lost = __ballot_sync(0xffffffffu, regen && !got);
            if (lane == 0u && lost) atomicSub(&S.active, __popc(lost));
            q_push(S.ready, alive || got, s, lane);
            q_push(S.freeq, regen && !got, s, lane);
            idle = 0;
        } else if (what == 2) {
            const int s = q_pop(S.miss, true, lane);
            if (s >= 0) {  // miss: sky / environment radiance * throughput into the accumulator, path ends
                const float4 ra = S.a[s], rb = S.b[s], rl = S.c[s];
                const float3 d = f3(rb);
                const float3 c = (LEGACY ? environment_color(sv, d) : sky_color(d)) * f3(rl);
                const unsigned pixel = __float_as_uint(ra.w);
                if (isfinite(c.x) && isfinite(c.y) && isfinite(c.z)) {
                    atomicAdd(&accum[pixel], make_float4(c.x, c.y, c.z, 1.0f));
                    if (rc.accum_sq) atomicAdd(&accum_sq[pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                }
            }
            const bool got = q_new_paths(S, U, s >= 0, s, rc, counters, lane, tiles_x, n_chunks, n_units, spp);
            const unsigned lost = __ballot_sync(0xffffffffu, s >= 0 && !got);
            if (lane == 0u && lost) atomicSub(&S.active, __popc(lost));
            q_push(S.ready, got, s, lane);
            q_push(S.freeq, s >= 0 && !got, s, lane);
            idle = 0;
        } else if (U.next < U.size && QUNI(S.freeq.count) > 0) {
            // this warp still holds camera paths of its work unit and other warps ran out: refill their freed slots
            const int s = q_pop(S.freeq, true, lane);
            const bool got = q_new_paths(S, U, s >= 0, s, rc, counters, lane, tiles_x, n_chunks, n_units, spp);
            const unsigned mg = __ballot_sync(0xffffffffu, got);
            if (lane == 0u && mg) atomicAdd(&S.active, __popc(mg));
            q_push(S.ready, got, s, lane);
            q_push(S.freeq, s >= 0 && !got, s, lane);
        } else if (n_empty == 32) {  // nothing to walk, nothing to shade: other warps hold the remaining paths
            if (QUNI(S.active) == 0 && U.next >= U.size) break;
            ++idle;
            __nanosleep(200);
        } else {
            idle = n_empty > 0 ? idle + 1 : 0;
        }
        __syncwarp();
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

int pt_render_queue(PtContext* ctx, const PtScene* s, const RenderConsts& rc, bool legacy, bool count, float4* accum,
                    float4* accum_sq, int serve_min) {
    cudaStream_t st = ctx->stream;
    int per_sm = 0;
    if (legacy) {
        if (count) PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_paths_queue<true, true>, Q_BLOCK, 0));
        else PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_paths_queue<true, false>, Q_BLOCK, 0));
    } else {
        if (count) PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_paths_queue<false, true>, Q_BLOCK, 0));
        else PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_paths_queue<false, false>, Q_BLOCK, 0));
    }
    if (per_sm < 1) per_sm = 1;
    int blocks = per_sm * ctx->sm_count;
    const unsigned long long need = (rc.total_paths + Q_SLOTS - 1) / Q_SLOTS;
    if ((unsigned long long)blocks > need) blocks = (int)need;
    if (blocks < 1) return PT_OK;
    if (legacy) {
        if (count) k_paths_queue<true, true><<<blocks, Q_BLOCK, 0, st>>>(s->view, rc, ctx->counters, accum, accum_sq, serve_min);
        else k_paths_queue<true, false><<<blocks, Q_BLOCK, 0, st>>>(s->view, rc, ctx->counters, accum, accum_sq, serve_min);
    } else {
        if (count) k_paths_queue<false, true><<<blocks, Q_BLOCK, 0, st>>>(s->view, rc, ctx->counters, accum, accum_sq, serve_min);
        else k_paths_queue<false, false><<<blocks, Q_BLOCK, 0, st>>>(s->view, rc, ctx->counters, accum, accum_sq, serve_min);
    }
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}
