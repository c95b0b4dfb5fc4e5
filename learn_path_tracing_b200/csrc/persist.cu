// persist.cu — persistent, ballot-scheduled kernels for scenes that have a BVH (PT_MODE_PERSIST).
//
// ncu on the K-step fused wavefront (k_paths) and on the one-ray-per-thread k_trace showed 6.0 / 6.7
// active threads per warp instruction on the mesh scenes and the 10M-triangle batch
// (profiles/r01_paths_yoimiya_v1_summary.txt, r01_trace_v1_summary.txt): a lane that misses everything
// waits for its neighbour's 100-node traversal, and lanes in a leaf wait for lanes in inner nodes.
// Here every LANE is its own small state machine (inner node / leaf / finished / idle) and the WARP votes
// by ballot on what to run next:
//
//   node phase     all lanes standing on an inner node take one node_step (two 256-bit loads, two slab
//                  tests); repeated while few lanes are waiting for anything else
//   service        entered once `serve_min` more lanes wait (at a leaf, finished, idle) than after the
//                  previous service, or nobody walks any more:
//     leaf phase     lanes on a leaf test their primitive and pop
//     shade phase    (render kernel, once >= shade_min lanes finished) miss -> RED.v4 into the
//                    accumulator, hit -> scatter; idle lanes take the next path of the warp's work unit
//                    and generate their camera ray in place (fused ray generation)
//     refill         (trace kernel) finished lanes write their hit record and draw the next ray
//
// A work unit of the render kernel is one 8x4 pixel tile x 4-16 samples (fewer in short launches), taken from ONE global counter
// (one atomic per 128-512 paths); the lanes of a warp therefore always look at the same 32 pixels, which
// keeps their rays — and their behaviour (all sky, all ground, all mesh) — alike.  One launch renders
// everything: no path pool in HBM, no per-bounce launches; HBM only sees the scene and the accumulator.
// The RNG is keyed on (pixel, sample, bounce), so the image is the same set of paths as the other modes.
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <string.h>

#include "wf_common.cuh"

#define ST_IDLE 0  // needs a new path
#define ST_TRAV 1  // traversal in progress (T.cur: inner node, leaf or PT_SENTINEL = finished)
#define ST_DEAD 2  // no work left for this lane
#define ST_NEW 3   // has a ray (camera or scattered), traversal not started yet

#define PT_TILE_W 8
#define PT_TILE_H 4
// samples per work unit: RenderConsts::unit_samples (chosen per launch in wavefront.cu:pt_render)

// WIDE (experimental, `make EXPERIMENTAL=1` + PT_WIDE=1 when the scene is built + PT_FLAG_WIDE): the node phase walks the
// 4-wide copy of the tree (extend.cuh:node_step4).  Measured on the B200 in round 2 (profiles/r02_ab_wide.txt): same
// images, 5-9 % SLOWER than the binary walk on every render workload, so it is not part of the default library.
// Shared-memory traversal stack (extend.cuh:TStack), entries per lane.  Same-box A/B on the B200
// (profiles/r02_ab_stack_lut_inline.txt): fixed-batch kernel 1649 (local stack) / 1673 (16 entries) / 1651 (24) / 1324
// (32: one block per SM fewer) Mrays/s; render kernels 10_final 5880 (local) / 5828 (8) / 5802 (16) / 5825 (24), Yoimiya
// within +-0.3 % — the stack traffic that fills half of the L1 sectors hits in L1 and is not what these kernels wait
// for, and the shared form costs two predicated instructions per push/pop.  Hence: 16 entries for the batch kernel,
// the plain local stack for the render kernels.
#ifndef PT_SSTACK_RENDER
#define PT_SSTACK_RENDER 0
#endif
#ifndef PT_SSTACK_TRACE
#define PT_SSTACK_TRACE 16
#endif
// KIND: 0 = v2 (taichi_pathtracer BSDFs), 1 = legacy textured meshes / spheres, 2 = legacy tutorial stages 6 / 7
// (untextured spheres, shade.cuh:scatter_legacy_stage) — its own instantiation, the two hot kernels do not carry it.
#define PT_KIND_V2 0
#define PT_KIND_LEGACY 1
#define PT_KIND_STAGE 2
template <int KIND, bool COUNT, bool WIDE = false>
__global__ void __launch_bounds__(PT_BLOCK, 4)
k_paths_persist(const SceneView sv, const RenderConsts rc, unsigned long long* __restrict__ counters,
                float4* __restrict__ accum, float4* __restrict__ accum_sq, int shade_min, int serve_min) {
    constexpr bool LEGACY = KIND == PT_KIND_LEGACY;
    constexpr int NS = WIDE ? 0 : PT_SSTACK_RENDER;
    __shared__ int s_stack[NS ? NS : 1][PT_BLOCK];
    int lstack[(WIDE ? PT_STACK_WIDE : PT_STACK) - NS];
    const TStack<NS> stack = {&s_stack[0][threadIdx.x], lstack};
#ifdef PT_OPT_GLOBAL_LUT
    const float* lut = sv.lut;
#else
    // texture transfer tables (albedo^2.2 | x^2 | 2x-1, 3 KB) in shared memory for the lifetime of the block
    __shared__ float s_lut[LEGACY ? 768 : 1];
    if (LEGACY && sv.lut) {
        for (int i = threadIdx.x; i < 768; i += PT_BLOCK) s_lut[i] = __ldg(sv.lut + i);
        __syncthreads();
    }
    const float* lut = s_lut;
#endif
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned tiles_x = ((unsigned)rc.W + PT_TILE_W - 1) / PT_TILE_W, tiles_y = ((unsigned)(rc.row1 - rc.row0) + PT_TILE_H - 1) / PT_TILE_H;
    const unsigned spp = rc.sample_end - rc.spp_offset;
    const unsigned n_chunks = (spp + rc.unit_samples - 1) / rc.unit_samples;
    const unsigned long long n_units = (unsigned long long)tiles_x * tiles_y * n_chunks;
    PathState p;
    Trav T;
    T.cur = PT_SENTINEL; T.sp = 1;
#ifdef PT_OPT_POSTPONE
    T.post = 0;
#endif
    int st = ST_IDLE;
    // warp-uniform: the work unit in progress and the next path of it
    unsigned unit_x0 = 0, unit_y0 = 0, unit_s0 = 0, unit_ns = 0, unit_next = 0, unit_size = 0;
    bool exhausted = false;  // the global unit counter has run past n_units
    int walk_min = 0;  // the node phase runs while more than walk_min lanes stand on an inner node
    unsigned nseg = 0;
    TraceCounters tc;
    tc.nodes = 0; tc.prims = 0;

    for (;;) {
        // invariant: T.cur == PT_SENTINEL unless st == ST_TRAV, so the vote only looks at T.cur
        const bool inner = PT_IS_INNER(T.cur);
        const int n_inner = __popc(__ballot_sync(0xffffffffu, inner));
        if (n_inner > walk_min) {
            if (inner) {
                if (WIDE) {
                    node_step4<COUNT>(sv, T, stack, tc);
                } else {
                    node_step<COUNT>(sv, T, stack, tc);
                    // a second step on the same vote (the vote costs about a quarter of a step)
                    if (n_inner > walk_min + 4 && PT_IS_INNER(T.cur)) node_step<COUNT>(sv, T, stack, tc);
                }
            }
            continue;
        }
        // ---- service ---------------------------------------------------------------------------------
        if (PT_TRAV_LEAFWORK(T)) leaf_step<COUNT>(sv, p.o, p.d, rc.tmin, T, stack, tc);
        __syncwarp();
        const unsigned m_fin = __ballot_sync(0xffffffffu, st == ST_TRAV && PT_TRAV_DONE(T));
        const unsigned m_idle = __ballot_sync(0xffffffffu, st == ST_IDLE);
        const unsigned m_walk = __ballot_sync(0xffffffffu, st == ST_TRAV && !PT_TRAV_DONE(T));
        if ((m_fin | m_idle) != 0u && (m_walk == 0u || (unsigned)__popc(m_fin | m_idle) >= (unsigned)shade_min)) {
            if (st == ST_TRAV && PT_TRAV_DONE(T)) {  // ---- shade
                if (T.h.prim < 0) {  // miss: sky / environment radiance * throughput, path ends
                    const float3 c = (LEGACY ? environment_color(sv, p.d) : sky_color(p.d)) * p.l;
                    if (isfinite(c.x) && isfinite(c.y) && isfinite(c.z)) {
                        atomicAdd(&accum[p.pixel], make_float4(c.x, c.y, c.z, 1.0f));
                        if (rc.accum_sq) atomicAdd(&accum_sq[p.pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                    }
                    st = ST_IDLE;
                } else if (!LEGACY && rc.shading_model == PT_SHADE_V2_NORMALS) {  // stages 4-5: normal as colour, no bounce
                    T.h.t = T.best;
                    const float3 c = normal_color(sv, p, T.h);
                    atomicAdd(&accum[p.pixel], make_float4(c.x, c.y, c.z, 1.0f));
                    if (rc.accum_sq) atomicAdd(&accum_sq[p.pixel], make_float4(c.x * c.x, c.y * c.y, c.z * c.z, 1.0f));
                    st = ST_IDLE;
                } else {
                    T.h.t = T.best;
                    if (LEGACY) scatter_legacy(sv, p, T.h, rc.absorptivity, rc.seed, lut);
                    else if (KIND == PT_KIND_STAGE) scatter_legacy_stage(sv, p, T.h, rc.shading_model, rc.absorptivity, rc.seed);
                    else scatter_v2(sv, p, T.h, rc.shading_model, rc.seed);
                    p.bounce += 1u;
                    st = p.bounce < (uint32_t)rc.max_depth ? ST_NEW : ST_IDLE;  // over propagate_limit: contributes nothing
                }
            }
            __syncwarp();
            // ---- refill idle lanes from the warp's work unit; take the next unit when it runs out
            for (;;) {
                const unsigned idle = __ballot_sync(0xffffffffu, st == ST_IDLE);
                if (idle == 0u) break;
                if (unit_next >= unit_size) {
                    if (exhausted) {
                        if (st == ST_IDLE) st = ST_DEAD;
                        break;
                    }
                    unsigned long long u = 0ull;
                    if (lane == 0u) u = atomicAdd(&counters[CNT_NEXT_PATH], 1ull);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= n_units) {
                        exhausted = true;
                        continue;
                    }
                    const unsigned tile = (unsigned)(u / n_chunks), chunk = (unsigned)(u - (unsigned long long)tile * n_chunks);
                    const unsigned ty = tile / tiles_x;
                    unit_x0 = (tile - ty * tiles_x) * PT_TILE_W;
                    unit_y0 = (unsigned)rc.row0 + ty * PT_TILE_H;
                    unit_s0 = chunk * rc.unit_samples;
                    unit_ns = min(rc.unit_samples, spp - unit_s0);
                    unit_size = unit_ns * 32u;
                    unit_next = 0u;
                }
                const unsigned take = min((unsigned)__popc(idle), unit_size - unit_next);
                if (st == ST_IDLE) {
                    const unsigned r = __popc(idle & lt);
                    if (r < take) {
                        const unsigned q = unit_next + r;
                        const unsigned px = unit_x0 + (q & 7u), py = unit_y0 + ((q >> 3) & 3u);
                        if (px < (unsigned)rc.W && py < (unsigned)rc.row1) {  // image sizes / row bands need not be tile multiples
                            p.pixel = py * (unsigned)rc.W + px;
                            p.sample = rc.spp_offset + unit_s0 + (q >> 5);
                            p.bounce = 0u;
                            p.l = f3(1.0f, 1.0f, 1.0f);
                            camera_ray(rc.cam, (int)px, (int)py, rng4(p.pixel, p.sample, 0u, rc.seed), &p.o, &p.d);
                            st = ST_NEW;
                        }
                    }
                }
                unit_next += take;
            }
            __syncwarp();
            if (st == ST_NEW) {  // continued and new paths start their next segment together
                trav_begin<COUNT>(sv, p.o, p.d, rc.tmin, INFINITY, T, stack, tc);
                ++nseg;
                st = ST_TRAV;
            }
            __syncwarp();
        }
        if (__ballot_sync(0xffffffffu, st != ST_DEAD) == 0u) break;  // every lane is dead
        // lanes that still wait (finished but not yet shaded, dead) do not count towards the next trigger
        walk_min = max(0, __popc(__ballot_sync(0xffffffffu, !PT_TRAV_DONE(T))) - serve_min);
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

// ------------------------------------------------------------------------------------------------
// Fixed ray batch (pt_trace_batch*, BASELINE configs[4]): persistent warps, rays drawn dynamically from
// one counter in the order given by `order` (a sort by entry point and direction, see k_ray_keys; NULL =
// batch order).  Result i is written to hits[i] of the ORIGINAL batch order.
// 6 blocks of 256 threads per SM (40 registers, a few spilled words): ncu shows this kernel latency-bound (long
// scoreboard 8.7 warps per issue at 4 blocks); measured 1.34 / 1.54 / 1.62 Grays/s at 4 / 5 / 6 blocks, 0.93 at 7+
// WIDE (experimental, opt-in): walks the 4-wide copy of the tree (extend.cuh:node_step4) with a 128-entry
// stack at 4 blocks/SM (three 256-bit loads per step want the registers).
template <bool COUNT, bool QNODES, bool WIDE = false>
__global__ void __launch_bounds__(PT_BLOCK, WIDE ? 4 : 6)
k_trace_persist(const SceneView sv, const float4* __restrict__ rays, const unsigned* __restrict__ order,
                float4* __restrict__ hits, long long n, unsigned long long* __restrict__ counters, int serve_min,
                int fetch_min) {
    constexpr int NS = WIDE ? 0 : PT_SSTACK_TRACE;
    __shared__ int s_stack[NS ? NS : 1][PT_BLOCK];
    int lstack[(WIDE ? PT_STACK_WIDE : PT_STACK) - NS];
    const TStack<NS> stack = {&s_stack[0][threadIdx.x], lstack};
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    float tmin = 0.0f;
    long long ray = -1;
    Trav T;
    T.cur = PT_SENTINEL; T.sp = 1;
#ifdef PT_OPT_POSTPONE
    T.post = 0;
#endif
    int st = ST_IDLE;
    bool exhausted = false;
    int walk_min = 0;
    TraceCounters tc;
    tc.nodes = 0; tc.prims = 0;
    for (;;) {
        const bool inner = PT_IS_INNER(T.cur);  // T.cur == PT_SENTINEL unless st == ST_TRAV
        const int n_inner = __popc(__ballot_sync(0xffffffffu, inner));
        if (n_inner > walk_min) {
            if (inner) {
                if (WIDE) node_step4<COUNT>(sv, T, stack, tc);
                else if (QNODES) node_step_q<COUNT>(sv, T, stack, tc);
                else node_step<COUNT>(sv, T, stack, tc);
                // a second step on the same vote (the vote costs about a quarter of a step)
                if (!WIDE && n_inner > walk_min + 4 && PT_IS_INNER(T.cur)) {
                    if (QNODES) node_step_q<COUNT>(sv, T, stack, tc);
                    else node_step<COUNT>(sv, T, stack, tc);
                }

            }
            continue;
        }
        // ---- service: leaves, then results + refill ------------------------------------------------------
        if (PT_TRAV_LEAFWORK(T)) leaf_step<COUNT>(sv, o, d, tmin, T, stack, tc);
        __syncwarp();
        const unsigned m_fin = __ballot_sync(0xffffffffu, st == ST_TRAV && PT_TRAV_DONE(T));
        const unsigned m_idle = __ballot_sync(0xffffffffu, st == ST_IDLE);
        const unsigned m_walk = __ballot_sync(0xffffffffu, st == ST_TRAV && !PT_TRAV_DONE(T));
        if ((m_fin | m_idle) != 0u && (m_walk == 0u || (unsigned)__popc(m_fin | m_idle) >= (unsigned)fetch_min)) {
            if (st == ST_TRAV && PT_TRAV_DONE(T)) {
                hits[ray] = make_float4(T.h.prim >= 0 ? T.best : -1.0f, __int_as_float(T.h.prim), T.h.u, T.h.v);
                st = ST_IDLE;
            }
            const unsigned idle = m_fin | m_idle;
            if (!exhausted) {
                const unsigned cnt = __popc(idle);
                unsigned long long base = 0ull;
                if (lane == 0u) base = atomicAdd(&counters[CNT_NEXT_PATH], (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                exhausted = (long long)(base + cnt) >= n;
                if (st == ST_IDLE) {
                    const long long k = (long long)base + __popc(idle & lt);
                    if (k < n) {
                        ray = order ? (long long)__ldg(&order[k]) : k;
                        const float4 ro = __ldg(&rays[2 * ray]), rd = __ldg(&rays[2 * ray + 1]);
                        o = f3(ro); d = f3(rd); tmin = ro.w;
                        trav_begin<COUNT>(sv, o, d, tmin, rd.w, T, stack, tc);
                        if (QNODES) trav_frame_q(sv, o, T);
                        st = ST_TRAV;
                    } else {
                        st = ST_DEAD;
                    }
                }
            } else if (st == ST_IDLE) {
                st = ST_DEAD;
            }
            __syncwarp();
        }
        if (__ballot_sync(0xffffffffu, st != ST_DEAD) == 0u) break;
        walk_min = max(0, __popc(__ballot_sync(0xffffffffu, !PT_TRAV_DONE(T))) - serve_min);
    }
    if (COUNT) {
        atomicAdd(&counters[CNT_NODES], (unsigned long long)tc.nodes);
        atomicAdd(&counters[CNT_PRIMS], (unsigned long long)tc.prims);
    }
}

// Sort key of a ray: Morton code over (entry point into the scene box: 3 x 6 bits, direction in the
// octahedral map: 2 x 6 bits), bits interleaved so that a run of consecutive keys is a thin beam.
PT_DEV unsigned spread5(unsigned v) {  // 6 bits -> every fifth bit
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) r |= ((v >> b) & 1u) << (5 * b);
    return r;
}
__global__ void k_ray_keys(const float4* __restrict__ rays, long long n, float3 lo, float3 inv_ext,
                           unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 ro = __ldg(&rays[2 * k]), rd = __ldg(&rays[2 * k + 1]);
    const float3 o = f3(ro), d = f3(rd);
    // entry point into the scene box (the origin when it is inside or the ray misses the box)
    const float3 inv = f3(1.0f / (fabsf(d.x) < 1e-18f ? copysignf(1e-18f, d.x) : d.x),
                          1.0f / (fabsf(d.y) < 1e-18f ? copysignf(1e-18f, d.y) : d.y),
                          1.0f / (fabsf(d.z) < 1e-18f ? copysignf(1e-18f, d.z) : d.z));
    const float3 hi = f3(lo.x + 1.0f / inv_ext.x, lo.y + 1.0f / inv_ext.y, lo.z + 1.0f / inv_ext.z);
    const float ax = (lo.x - o.x) * inv.x, bx = (hi.x - o.x) * inv.x;
    const float ay = (lo.y - o.y) * inv.y, by = (hi.y - o.y) * inv.y;
    const float az = (lo.z - o.z) * inv.z, bz = (hi.z - o.z) * inv.z;
    const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    const float te = t0 <= t1 ? t0 : 0.0f;
    const float px = (o.x + te * d.x - lo.x) * inv_ext.x, py = (o.y + te * d.y - lo.y) * inv_ext.y,
                pz = (o.z + te * d.z - lo.z) * inv_ext.z;
    // octahedral map of the direction -> [0,1)^2
    const float s = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
    float ux = d.x * s, uy = d.y * s;
    if (d.z < 0.0f) {
        const float tx = (1.0f - fabsf(uy)) * (ux >= 0.0f ? 1.0f : -1.0f);
        const float ty = (1.0f - fabsf(ux)) * (uy >= 0.0f ? 1.0f : -1.0f);
        ux = tx; uy = ty;
    }
    const float S = 64.0f;
    const unsigned qx = (unsigned)fminf(fmaxf(px * S, 0.0f), S - 1.0f), qy = (unsigned)fminf(fmaxf(py * S, 0.0f), S - 1.0f),
                   qz = (unsigned)fminf(fmaxf(pz * S, 0.0f), S - 1.0f);
    const unsigned qu = (unsigned)fminf(fmaxf((0.5f * ux + 0.5f) * S, 0.0f), S - 1.0f),
                   qv = (unsigned)fminf(fmaxf((0.5f * uy + 0.5f) * S, 0.0f), S - 1.0f);
    keys[k] = spread5(qx) << 4 | spread5(qy) << 3 | spread5(qz) << 2 | spread5(qu) << 1 | spread5(qv);
    vals[k] = (unsigned)k;
}

static int ensure_sort_scratch(PtContext* ctx, size_t bytes) {
    if (bytes <= ctx->sort_bytes) return PT_OK;
    if (ctx->sort_scratch) cudaFree(ctx->sort_scratch);
    ctx->sort_scratch = nullptr;
    ctx->sort_bytes = 0;
    PT_CUDA(cudaMalloc(&ctx->sort_scratch, bytes));
    ctx->sort_bytes = bytes;
    return PT_OK;
}

// Persistent grid: as many blocks as are resident at once.
template <class K>
static int resident_blocks(PtContext* ctx, K kernel, int* out) {
    // Shared memory and L1 share 256 KB per SM: ask for exactly the carveout that the blocks the register budget allows
    // (__launch_bounds__: 4 or 6 per SM) need for their stacks, the rest stays L1 for nodes and triangles.
    static_assert(PT_SSTACK_TRACE * PT_BLOCK * 4 * 6 + 6 * 1024 <= 227 * 1024, "trace stacks exceed the shared memory of an SM");
    static_assert(PT_SSTACK_RENDER * PT_BLOCK * 4 * 4 + 4 * 1024 <= 227 * 1024, "render stacks exceed the shared memory of an SM");
    cudaFuncAttributes fa;
    PT_CUDA(cudaFuncGetAttributes(&fa, kernel));
    if (fa.sharedSizeBytes > 1024) {
        const int by_regs = fa.numRegs > 0 ? 65536 / (fa.numRegs * PT_BLOCK) : 1;
        const size_t want = (size_t)(by_regs < 1 ? 1 : by_regs) * (fa.sharedSizeBytes + 1024);
        int pct = (int)((want * 100 + 228 * 1024 - 1) / (228 * 1024));
        PT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
    }
    int per_sm = 0;
    PT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PT_BLOCK, 0));
    if (per_sm < 1) per_sm = 1;
    *out = per_sm * ctx->sm_count;
    return PT_OK;
}

int pt_trace_persist(PtContext* ctx, const PtScene* s, const float4* rays, long long n, float4* hits, bool count,
                     bool sort, bool use_qnodes, int serve_min, int fetch_min, cudaEvent_t ev_sorted, bool wide) {
    cudaStream_t st = ctx->stream;
    const unsigned* order = nullptr;
    if (sort && n > 1) {
        PT_REQUIRE(n < (1ll << 31), "sorted ray batches are limited to 2^31 - 1 rays (cub::DeviceRadixSort takes an int count)");
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<unsigned> kb(nullptr, nullptr), vb(nullptr, nullptr);
        PT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 30, st));
        const size_t arr = ((size_t)n * sizeof(unsigned) + 255) / 256 * 256;
        int rcs = ensure_sort_scratch(ctx, 4 * arr + tmp_bytes);
        if (rcs) return rcs;
        char* base = (char*)ctx->sort_scratch;
        kb = cub::DoubleBuffer<unsigned>((unsigned*)base, (unsigned*)(base + arr));
        vb = cub::DoubleBuffer<unsigned>((unsigned*)(base + 2 * arr), (unsigned*)(base + 3 * arr));
        const float3 lo = make_float3(s->bounds_lo[0], s->bounds_lo[1], s->bounds_lo[2]);
        const float3 ie = make_float3(1.0f / fmaxf(s->bounds_hi[0] - s->bounds_lo[0], 1e-30f),
                                      1.0f / fmaxf(s->bounds_hi[1] - s->bounds_lo[1], 1e-30f),
                                      1.0f / fmaxf(s->bounds_hi[2] - s->bounds_lo[2], 1e-30f));
        k_ray_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rays, n, lo, ie, kb.Current(), vb.Current());
        PT_CUDA(cub::DeviceRadixSort::SortPairs(base + 4 * arr, tmp_bytes, kb, vb, (int)n, 0, 30, st));
        order = vb.Current();
    }
    if (ev_sorted) PT_CUDA(cudaEventRecord(ev_sorted, st));
#ifdef PT_EXPERIMENTAL
    if (wide) {  // experimental 4-wide walk (PT_FLAG_TRACE_WIDE): float nodes only
        int wblocks = 0;
        int rcw = count ? resident_blocks(ctx, k_trace_persist<true, false, true>, &wblocks) : resident_blocks(ctx, k_trace_persist<false, false, true>, &wblocks);
        if (rcw) return rcw;
        const long long wneed = (n + PT_BLOCK - 1) / PT_BLOCK;
        if (wblocks > wneed) wblocks = (int)wneed;
        if (count) k_trace_persist<true, false, true><<<wblocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
        else k_trace_persist<false, false, true><<<wblocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
        PT_CUDA(cudaGetLastError());
        return PT_OK;
    }
#else
    PT_REQUIRE(!wide, "the 4-wide walk is an experimental kernel form: rebuild with `make EXPERIMENTAL=1`");
#endif
    int blocks = 0, rcb;
    const bool q = s->view.qnodes != nullptr && use_qnodes;
    if (count) rcb = q ? resident_blocks(ctx, k_trace_persist<true, true>, &blocks) : resident_blocks(ctx, k_trace_persist<true, false>, &blocks);
    else rcb = q ? resident_blocks(ctx, k_trace_persist<false, true>, &blocks) : resident_blocks(ctx, k_trace_persist<false, false>, &blocks);
    if (rcb) return rcb;
    const long long need = (n + PT_BLOCK - 1) / PT_BLOCK;
    if (blocks > need) blocks = (int)need;
    if (count) {
        if (q) k_trace_persist<true, true><<<blocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
        else k_trace_persist<true, false><<<blocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
    } else {
        if (q) k_trace_persist<false, true><<<blocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
        else k_trace_persist<false, false><<<blocks, PT_BLOCK, 0, st>>>(s->view, rays, order, hits, n, ctx->counters, serve_min, fetch_min);
    }
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

template <int KIND, bool COUNT, bool WIDE>
static int launch_persist(PtContext* ctx, const PtScene* s, const RenderConsts& rc, float4* accum, float4* accum_sq, int shade_min,
                          int serve_min) {
    int blocks = 0;
    int rcb = resident_blocks(ctx, k_paths_persist<KIND, COUNT, WIDE>, &blocks);
    if (rcb) return rcb;
    const unsigned long long need = (rc.total_paths + PT_BLOCK - 1) / PT_BLOCK;
    if ((unsigned long long)blocks > need) blocks = (int)need;
    if (blocks < 1) return PT_OK;
    k_paths_persist<KIND, COUNT, WIDE><<<blocks, PT_BLOCK, 0, ctx->stream>>>(s->view, rc, ctx->counters, accum, accum_sq, shade_min, serve_min);
    PT_CUDA(cudaGetLastError());
    return PT_OK;
}

int pt_render_persist(PtContext* ctx, const PtScene* s, const RenderConsts& rc, bool legacy, bool count, float4* accum,
                      float4* accum_sq, int shade_min, int serve_min, bool wide) {
#ifdef PT_EXPERIMENTAL
    if (wide) {  // experimental: the node phase walks the 4-wide copy of the tree
        if (legacy) return count ? launch_persist<PT_KIND_LEGACY, true, true>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
                                 : launch_persist<PT_KIND_LEGACY, false, true>(ctx, s, rc, accum, accum_sq, shade_min, serve_min);
        return count ? launch_persist<PT_KIND_V2, true, true>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
                     : launch_persist<PT_KIND_V2, false, true>(ctx, s, rc, accum, accum_sq, shade_min, serve_min);
    }
#else
    PT_REQUIRE(!wide, "the 4-wide walk is an experimental kernel form: rebuild with `make EXPERIMENTAL=1`");
#endif
    if (rc.shading_model == PT_SHADE_LEGACY_STAGE6 || rc.shading_model == PT_SHADE_LEGACY_STAGE7)
        return count ? launch_persist<PT_KIND_STAGE, true, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
                     : launch_persist<PT_KIND_STAGE, false, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min);
    if (legacy) return count ? launch_persist<PT_KIND_LEGACY, true, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
                             : launch_persist<PT_KIND_LEGACY, false, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min);
    return count ? launch_persist<PT_KIND_V2, true, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min)
                 : launch_persist<PT_KIND_V2, false, false>(ctx, s, rc, accum, accum_sq, shade_min, serve_min);
}
