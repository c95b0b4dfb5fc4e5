// lbvh.cu — linear BVH built entirely on the GPU (no reference equivalent on the device: the
// reference builds a SAH tree in host Python, legacy/PT_in_one_weekend/15_module.py:608-634,716-754).
//
//   1. centroid bounds                (block reduce + ordered-int atomics)
//   2. 63-bit Morton codes            (21 bits per axis)
//   3. radix sort of (code, prim)     (cub::DeviceRadixSort — library infrastructure, not the hot path)
//   4. Karras 2012 hierarchy          (one thread per internal node, duplicate codes broken by index)
//   5. bottom-up refit + emit         (second arrival at a node owns it; writes the 64-byte BVH2 node
//                                      that stores BOTH child boxes, the layout traversal reads)
#include <cub/device/device_radix_sort.cuh>

#include "pt_internal.h"

namespace {

__device__ __forceinline__ int f2ord(float f) {  // monotone float -> int
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_centroid_bounds(const float4* __restrict__ aabb, const int* __restrict__ ids, long long n, int* bounds) {
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int p = ids[k];
        const float4 a = aabb[2 * (size_t)p], b = aabb[2 * (size_t)p + 1];
        const float cx = 0.5f * (a.x + b.x), cy = 0.5f * (a.y + b.y), cz = 0.5f * (a.z + b.z);
        lo.x = fminf(lo.x, cx); lo.y = fminf(lo.y, cy); lo.z = fminf(lo.z, cz);
        hi.x = fmaxf(hi.x, cx); hi.y = fmaxf(hi.y, cy); hi.z = fmaxf(hi.z, cz);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bounds[0], f2ord(lo.x)); atomicMin(&bounds[1], f2ord(lo.y)); atomicMin(&bounds[2], f2ord(lo.z));
        atomicMax(&bounds[3], f2ord(hi.x)); atomicMax(&bounds[4], f2ord(hi.y)); atomicMax(&bounds[5], f2ord(hi.z));
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float4* __restrict__ aabb, const int* __restrict__ ids, long long n, const int* __restrict__ bounds,
                         unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float lx = ord2f(bounds[0]), ly = ord2f(bounds[1]), lz = ord2f(bounds[2]);
    const float sx = ord2f(bounds[3]) - lx, sy = ord2f(bounds[4]) - ly, sz = ord2f(bounds[5]) - lz;
    const int p = ids[k];
    const float4 a = aabb[2 * (size_t)p], b = aabb[2 * (size_t)p + 1];
    const float cx = 0.5f * (a.x + b.x), cy = 0.5f * (a.y + b.y), cz = 0.5f * (a.z + b.z);
    const float S = 2097152.0f;  // 2^21
    const float qx = sx > 0.0f ? fminf(fmaxf((cx - lx) / sx * S, 0.0f), S - 1.0f) : 0.0f;
    const float qy = sy > 0.0f ? fminf(fmaxf((cy - ly) / sy * S, 0.0f), S - 1.0f) : 0.0f;
    const float qz = sz > 0.0f ? fminf(fmaxf((cz - lz) / sz * S, 0.0f), S - 1.0f) : 0.0f;
    keys[k] = expand21((unsigned long long)qx) << 2 | expand21((unsigned long long)qy) << 1 | expand21((unsigned long long)qz);
    vals[k] = p;
}

// common-prefix length of sorted keys i and j; ties broken by index (Karras 2012, section 4)
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, long long n, long long i, long long j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clzll((unsigned long long)(i ^ j));
    return __clzll(a ^ b);
}

// child encoding during the build: >= 0 internal node, < 0 leaf ~sorted_index
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, long long n, int2* __restrict__ children,
                            int* __restrict__ parent) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    long long lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    long long l = 0;
    for (long long t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const long long j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    long long s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const long long gamma = i + s * d + (d < 0 ? -1 : 0);
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    int left, right;
    if (lo == gamma) { left = ~(int)gamma; parent[(n - 1) + gamma] = (int)i; }
    else { left = (int)gamma; parent[gamma] = (int)i; }
    if (hi == gamma + 1) { right = ~(int)(gamma + 1); parent[(n - 1) + gamma + 1] = (int)i; }
    else { right = (int)(gamma + 1); parent[gamma + 1] = (int)i; }
    children[i] = make_int2(left, right);
    if (i == 0) parent[0] = -1;
}

__device__ __forceinline__ void child_box(int c, const int* __restrict__ sorted_prim, const float4* __restrict__ aabb,
                                          const float4* node_box, float4* lo, float4* hi) {
    if (c < 0) {
        const int p = sorted_prim[~c];
        *lo = aabb[2 * (size_t)p];
        *hi = aabb[2 * (size_t)p + 1];
    } else {  // written by another thread earlier in this kernel: bypass L1
        *lo = __ldcg(&node_box[2 * (size_t)c]);
        *hi = __ldcg(&node_box[2 * (size_t)c + 1]);
    }
}

__global__ void k_refit_emit(long long n, const int2* __restrict__ children, const int* __restrict__ parent,
                             const int* __restrict__ sorted_prim, const float4* __restrict__ aabb, float4* node_box,
                             int* flags, float4* __restrict__ nodes) {
    const long long leaf = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int cur = parent[(n - 1) + leaf];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1) == 0) return;  // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const int2 ch = children[cur];
        float4 l0, h0, l1, h1;
        child_box(ch.x, sorted_prim, aabb, node_box, &l0, &h0);
        child_box(ch.y, sorted_prim, aabb, node_box, &l1, &h1);
        const int r0 = ch.x < 0 ? ~sorted_prim[~ch.x] : ch.x;
        const int r1 = ch.y < 0 ? ~sorted_prim[~ch.y] : ch.y;
        float4* o = nodes + 4 * (size_t)cur;
        o[0] = make_float4(l0.x, l0.y, l0.z, h0.x);
        o[1] = make_float4(h0.y, h0.z, l1.x, l1.y);
        o[2] = make_float4(l1.z, h1.x, h1.y, h1.z);
        o[3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
        __stcg(&node_box[2 * (size_t)cur], make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.0f));
        __stcg(&node_box[2 * (size_t)cur + 1], make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.0f));
        __threadfence();
        cur = parent[cur];
    }
}

// Quantised copy of the nodes for big trees: child boxes as u16 in the root frame, conservative (a min never
// dequantises above the float min, a max never below the float max, checked with the traversal's own fmaf).
__device__ __forceinline__ unsigned q_down(float v, float lo, float scale) {
    int q = (int)floor(((double)v - (double)lo) / (double)scale);
    q = q < 0 ? 0 : (q > 65535 ? 65535 : q);
    while (q > 0 && fmaf((float)q, scale, lo) > v) --q;
    return (unsigned)q;
}
__device__ __forceinline__ unsigned q_up(float v, float lo, float scale) {
    int q = (int)ceil(((double)v - (double)lo) / (double)scale);
    q = q < 0 ? 0 : (q > 65535 ? 65535 : q);
    while (q < 65535 && fmaf((float)q, scale, lo) < v) ++q;
    return (unsigned)q;
}
__global__ void k_quantize_nodes(const float4* __restrict__ nodes, long long n, float3 lo, float3 sc, uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = nodes[4 * i], b = nodes[4 * i + 1], c = nodes[4 * i + 2], k = nodes[4 * i + 3];
    // child 0: min a.x a.y a.z max a.w b.x b.y; child 1: min b.z b.w c.x max c.y c.z c.w
    const unsigned m0x = q_down(a.x, lo.x, sc.x), m0y = q_down(a.y, lo.y, sc.y), m0z = q_down(a.z, lo.z, sc.z);
    const unsigned M0x = q_up(a.w, lo.x, sc.x), M0y = q_up(b.x, lo.y, sc.y), M0z = q_up(b.y, lo.z, sc.z);
    const unsigned m1x = q_down(b.z, lo.x, sc.x), m1y = q_down(b.w, lo.y, sc.y), m1z = q_down(c.x, lo.z, sc.z);
    const unsigned M1x = q_up(c.y, lo.x, sc.x), M1y = q_up(c.z, lo.y, sc.y), M1z = q_up(c.w, lo.z, sc.z);
    out[2 * i] = make_uint4(m0x | m0y << 16, m0z | M0x << 16, M0y | M0z << 16, m1x | m1y << 16);
    out[2 * i + 1] = make_uint4(m1z | M1x << 16, M1y | M1z << 16, __float_as_uint(k.x), __float_as_uint(k.y));
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t alloc(T** p, size_t n) {
        cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace

int pt_lbvh_build(PtContext* ctx, const float4* d_prim_aabb, const int* d_local_ids, int64_t n, float4** d_nodes_out,
                  int64_t* n_nodes_out, int* root_out) {
    PT_REQUIRE(n >= 2 && n < (1ll << 31), "LBVH needs 2 <= n < 2^31 primitives");
    cudaStream_t st = ctx->stream;
    Scratch sc;
    int* bounds = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int *vals = nullptr, *vals2 = nullptr, *parent = nullptr, *flags = nullptr;
    int2* children = nullptr;
    float4* node_box = nullptr;
    PT_CUDA(sc.alloc(&bounds, 6));
    PT_CUDA(sc.alloc(&keys, (size_t)n));
    PT_CUDA(sc.alloc(&keys2, (size_t)n));
    PT_CUDA(sc.alloc(&vals, (size_t)n));
    PT_CUDA(sc.alloc(&vals2, (size_t)n));
    PT_CUDA(sc.alloc(&parent, (size_t)(2 * n - 1)));
    PT_CUDA(sc.alloc(&flags, (size_t)(n - 1)));
    PT_CUDA(sc.alloc(&children, (size_t)(n - 1)));
    PT_CUDA(sc.alloc(&node_box, (size_t)(2 * (n - 1))));
    float4* nodes = nullptr;
    PT_CUDA(cudaMalloc(&nodes, (size_t)(n - 1) * 4 * sizeof(float4)));

    const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    PT_CUDA(cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, st));
    const unsigned B = 256;
    const unsigned gridN = (unsigned)((n + B - 1) / B);
    k_centroid_bounds<<<gridN < 1184u ? gridN : 1184u, B, 0, st>>>(d_prim_aabb, d_local_ids, n, bounds);
    k_morton<<<gridN, B, 0, st>>>(d_prim_aabb, d_local_ids, n, bounds, keys, vals);

    size_t tmp_bytes = 0;
    cub::DoubleBuffer<unsigned long long> kb(keys, keys2);
    cub::DoubleBuffer<int> vb(vals, vals2);
    PT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 63, st));
    void* tmp = nullptr;
    PT_CUDA(sc.alloc((char**)&tmp, tmp_bytes));
    PT_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, 63, st));

    PT_CUDA(cudaMemsetAsync(flags, 0, (size_t)(n - 1) * sizeof(int), st));
    k_hierarchy<<<(unsigned)((n - 1 + B - 1) / B), B, 0, st>>>(kb.Current(), n, children, parent);
    k_refit_emit<<<gridN, B, 0, st>>>(n, children, parent, vb.Current(), d_prim_aabb, node_box, flags, nodes);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(nodes);
        pt_set_error("pt_lbvh_build: %s", cudaGetErrorString(e));
        return PT_ERR_CUDA;
    }
    *d_nodes_out = nodes;
    *n_nodes_out = n - 1;
    *root_out = 0;
    return PT_OK;
}

int pt_quantize_nodes(PtContext* ctx, const float4* d_nodes, int64_t n_nodes, const float lo[3], const float scale[3],
                      uint4** d_qnodes_out) {
    uint4* q = nullptr;
    PT_CUDA(cudaMalloc(&q, (size_t)n_nodes * 2 * sizeof(uint4)));
    k_quantize_nodes<<<(unsigned)((n_nodes + 255) / 256), 256, 0, ctx->stream>>>(
        d_nodes, n_nodes, make_float3(lo[0], lo[1], lo[2]), make_float3(scale[0], scale[1], scale[2]), q);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(q);
        pt_set_error("pt_quantize_nodes: %s", cudaGetErrorString(e));
        return PT_ERR_CUDA;
    }
    *d_qnodes_out = q;
    return PT_OK;
}
