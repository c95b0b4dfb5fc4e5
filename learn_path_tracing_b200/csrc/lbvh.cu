// lbvh.cu — linear BVH built entirely on the GPU (no reference equivalent on the device: the
// reference builds a SAH tree in host Python, legacy/PT_in_one_weekend/15_module.py:608-634,716-754).
//
//   1. centroid bounds                (block reduce + ordered-int atomics)
//   2. 63-bit Morton codes            (21 bits per axis)
//   3. radix sort of (code, prim)     (cub::DeviceRadixSort — library infrastructure, not the hot path)
//   4. Karras 2012 hierarchy          (one thread per internal node, duplicate codes broken by index)
//   5. bottom-up refit + emit         (second arrival at a node owns it; writes the 64-byte BVH2 node
//                                      that stores BOTH child boxes, the layout traversal reads)
//
// For trees of up to PT_PLOC_MAX primitives step 4 is replaced by a locally-ordered agglomerative clustering over the
// same Morton order (PLOC, Meister & Bittner 2018): every cluster looks PT_PLOC_RADIUS neighbours to either side for
// the partner with the smallest merged surface area, mutual choices merge, the survivors are compacted, until one
// cluster is left.  Same inputs and outputs as k_hierarchy (children / parent arrays), a better tree for the same
// sort; the closest hit does not depend on the shape of the tree, so every parity test holds for both.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "pt_internal.h"

#define PT_MAX_TREE_DEPTH 60   // candidates deeper than this are discarded (traversal stack: 64 entries); the Karras tree
                               // is bounded by its 63 key bits + log2 of the longest run of equal codes
#define PT_HOST_SAH_MAX 4096   // up to here a full-sweep SAH tree from the host competes as well (see host_sah_rec)
#define PT_PLOC_MAX (1 << 20)  // larger trees keep the Karras hierarchy (one pass instead of ~30 rounds over all clusters)

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
                             int* flags, float4* __restrict__ nodes, int* subtree /* inner nodes per subtree, or NULL */) {
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
        if (subtree) __stcg(&subtree[cur], 1 + (ch.x >= 0 ? __ldcg(&subtree[ch.x]) : 0) + (ch.y >= 0 ? __ldcg(&subtree[ch.y]) : 0));
        __threadfence();
        cur = parent[cur];
    }
}

// ---- PLOC ----------------------------------------------------------------------------------------
#define PT_PLOC_RADIUS 16

__global__ void k_ploc_init(long long n, const int* __restrict__ sorted_prim, const float4* __restrict__ aabb,
                            int* __restrict__ ref, float4* __restrict__ lo, float4* __restrict__ hi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = sorted_prim[i];
    ref[i] = ~(int)i;  // leaf: sorted index, as k_hierarchy encodes it
    lo[i] = aabb[2 * (size_t)p];
    hi[i] = aabb[2 * (size_t)p + 1];
}

__device__ __forceinline__ float merged_area(float4 al, float4 ah, float4 bl, float4 bh) {
    const float x = fmaxf(ah.x, bh.x) - fminf(al.x, bl.x), y = fmaxf(ah.y, bh.y) - fminf(al.y, bl.y),
                z = fmaxf(ah.z, bh.z) - fminf(al.z, bl.z);
    return x * y + y * z + z * x;
}

// nearest neighbour of every cluster within the radius: smallest merged area; ties go to the even/odd partner i^1 if it
// is among them (a run of identical boxes — duplicated faces — then pairs up and halves every round instead of merging
// once per round), else to the lowest index.  With that rule the lexicographically first of the globally closest pairs
// is always mutual, so every round merges at least once.
__global__ void k_ploc_nn(int c, int radius, const float4* __restrict__ lo, const float4* __restrict__ hi, int* __restrict__ nn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float4 al = lo[i], ah = hi[i];
    float best = INFINITY;
    int bj = -1;
    const int j0 = max(0, i - radius), j1 = min(c - 1, i + radius);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const float a = merged_area(al, ah, lo[j], hi[j]);
        if (a < best || (a == best && j == (i ^ 1))) { best = a; bj = j; }
    }
    nn[i] = bj;
}

// mutual pairs merge into a new node (indices are handed out downwards: the very last merge, alone in its round,
// gets 0 = the root); the lower partner carries the merged cluster on, the upper one drops out
__global__ void k_ploc_merge(int c, long long n, const int* __restrict__ nn, int* __restrict__ ref, float4* __restrict__ lo,
                             float4* __restrict__ hi, int* __restrict__ keep, int* __restrict__ next_node,
                             int2* __restrict__ children, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const int j = nn[i];
    if (j < 0 || nn[j] != i) { keep[i] = 1; return; }
    if (i > j) { keep[i] = 0; return; }
    const int idx = atomicSub(next_node, 1);
    const int ri = ref[i], rj = ref[j];
    children[idx] = make_int2(ri, rj);
    parent[ri < 0 ? (n - 1) + ~ri : ri] = idx;
    parent[rj < 0 ? (n - 1) + ~rj : rj] = idx;
    const float4 al = lo[i], ah = hi[i], bl = lo[j], bh = hi[j];
    // nobody else reads lo/hi/ref of cluster i in this kernel (the neighbour search is over)
    lo[i] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.0f);
    hi[i] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.0f);
    ref[i] = idx;
    keep[i] = 1;
}

__global__ void k_ploc_compact(int c, const int* __restrict__ keep, const int* __restrict__ pos, const int* __restrict__ ref,
                               const float4* __restrict__ lo, const float4* __restrict__ hi, int* __restrict__ ref2,
                               float4* __restrict__ lo2, float4* __restrict__ hi2, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    if (keep[i]) {
        const int q = pos[i];
        ref2[q] = ref[i]; lo2[q] = lo[i]; hi2[q] = hi[i];
    }
    if (i == c - 1) *count = pos[i] + keep[i];
}

// Depth-first (pre-order) renumbering of a tree whose nodes were numbered in creation order (PLOC: by merge round, so
// a parent sits far from its children): the position of a node is the number of inner nodes visited before it, read
// off the path to the root from the subtree sizes.  Afterwards a subtree is one contiguous run of nodes and the first
// child follows its parent directly — half of the descents stay in the 128-byte line they came from.
__global__ void k_dfs_index(long long n_nodes, const int2* __restrict__ children, const int* __restrict__ parent,
                            const int* __restrict__ subtree, int* __restrict__ newid, int* __restrict__ max_depth) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    int idx = 0, c = (int)i, p, depth = 2;  // this node + the leaves below it
    while ((p = parent[c]) >= 0) {
        const int2 ch = children[p];
        idx += 1 + ((ch.y == c && ch.x >= 0) ? subtree[ch.x] : 0);
        c = p;
        ++depth;
    }
    newid[i] = idx;
    if (depth > PT_MAX_TREE_DEPTH) atomicMax(max_depth, depth);
}
__global__ void k_relayout(long long n_nodes, const float4* __restrict__ in, const int* __restrict__ newid, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const float4 a = in[4 * i], b = in[4 * i + 1], c = in[4 * i + 2];
    float4 k = in[4 * i + 3];
    const int r0 = __float_as_int(k.x), r1 = __float_as_int(k.y);
    if (r0 >= 0) k.x = __int_as_float(newid[r0]);
    if (r1 >= 0) k.y = __int_as_float(newid[r1]);
    float4* o = out + 4 * (size_t)newid[i];
    o[0] = a; o[1] = b; o[2] = c; o[3] = k;
}

// Depth of a hierarchy (levels of inner nodes above the deepest leaf) from its parent array, one thread per leaf.  Every
// candidate is measured: the traversal stacks hold PT_STACK entries and are not bounds-checked in the kernels.
__global__ void k_tree_depth(long long n, const int* __restrict__ parent, int* __restrict__ max_depth) {
    const long long leaf = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int depth = 0;
    for (int c = parent[(n - 1) + leaf]; c >= 0; c = parent[c]) ++depth;
    for (int o = 16; o > 0; o >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, depth);
}
// keys[i] = i: the Karras hierarchy over these is the binary radix tree of the sorted INDEX, at most 32 levels deep
// whatever the geometry — the fallback when every real candidate is too deep for the traversal stack
__global__ void k_index_keys(long long n, unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (unsigned long long)i;
}

// SAH cost of an emitted tree up to constants: the sum of the surface areas of every child box
__global__ void k_sah_cost(const float4* __restrict__ nodes, long long n_nodes, double* __restrict__ cost) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0;
    if (i < n_nodes) {
        const float4 p = nodes[4 * i], q = nodes[4 * i + 1], r = nodes[4 * i + 2];
        const float x0 = p.w - p.x, y0 = q.x - p.y, z0 = q.y - p.z;   // child 0: min p.xyz max p.w q.x q.y
        const float x1 = r.y - q.z, y1 = r.z - q.w, z1 = r.w - r.x;   // child 1: min q.z q.w r.x max r.yzw
        a = (double)(x0 * y0 + y0 * z0 + z0 * x0) + (double)(x1 * y1 + y1 * z1 + z1 * x1);
    }
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0 && a != 0.0) atomicAdd(cost, a);
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

// ---- host full-sweep SAH for tiny trees ------------------------------------------------------------
// A few hundred primitives (the RTIOW sphere field: 485) are not worth a device builder's launches, and neither Morton
// hierarchy is good on them (CPU prototype, tools/tree_quality_proto.py: full-sweep SAH cost 4561 against 5162 Karras /
// 5378 PLOC, 3-10 % fewer node visits).  Top-down, every split evaluated on all three axes, one primitive per leaf,
// nodes emitted depth-first (root = 0) in the layout k_refit_emit writes.  O(n log^2 n): ~0.2 ms for 485 primitives.
struct HostBox { float lo[3], hi[3]; };
static inline float hb_area(const HostBox& b) {
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}
static inline void hb_grow(HostBox& a, const HostBox& b) {
    for (int c = 0; c < 3; ++c) { a.lo[c] = std::min(a.lo[c], b.lo[c]); a.hi[c] = std::max(a.hi[c], b.hi[c]); }
}
// builds the subtree over idx[begin, end) (indices into box/prim); returns its child reference, its box in *out
static int host_sah_rec(std::vector<int>& idx, int begin, int end, const std::vector<HostBox>& box, const std::vector<int>& prim,
                        std::vector<float>& nodes, std::vector<float>& tmp_area, HostBox* out, int depth, int* max_depth) {
    const int cnt = end - begin;
    if (depth > *max_depth) *max_depth = depth;
    if (cnt == 1) { *out = box[idx[begin]]; return ~prim[idx[begin]]; }
    int best_axis = 0, best_k = 1;
    float best_cost = INFINITY;
    std::vector<int> order(idx.begin() + begin, idx.begin() + end), best_order;
    for (int ax = 0; ax < 3; ++ax) {
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return box[a].lo[ax] + box[a].hi[ax] < box[b].lo[ax] + box[b].hi[ax];
        });
        HostBox acc = box[order[cnt - 1]];
        tmp_area[cnt - 1] = hb_area(acc);
        for (int k = cnt - 2; k >= 1; --k) { hb_grow(acc, box[order[k]]); tmp_area[k] = hb_area(acc); }  // area of [k, cnt)
        acc = box[order[0]];
        for (int k = 1; k < cnt; ++k) {   // split: [0, k) | [k, cnt)
            const float cost = hb_area(acc) * (float)k + tmp_area[k] * (float)(cnt - k);
            // equal costs (identical boxes): the more balanced split, or the tree degenerates into a chain
            if (cost < best_cost || (cost == best_cost && abs(2 * k - cnt) < abs(2 * best_k - cnt))) {
                best_cost = cost; best_axis = ax; best_k = k;
            }
            hb_grow(acc, box[order[k]]);
        }
        if (best_axis == ax) best_order = order;
    }
    std::copy(best_order.begin(), best_order.end(), idx.begin() + begin);
    const size_t me = nodes.size() / 16;
    nodes.resize(nodes.size() + 16, 0.0f);
    HostBox b0, b1;
    const int r0 = host_sah_rec(idx, begin, begin + best_k, box, prim, nodes, tmp_area, &b0, depth + 1, max_depth);
    const int r1 = host_sah_rec(idx, begin + best_k, end, box, prim, nodes, tmp_area, &b1, depth + 1, max_depth);
    float* o = &nodes[16 * me];
    for (int c = 0; c < 3; ++c) { o[c] = b0.lo[c]; o[3 + c] = b0.hi[c]; o[6 + c] = b1.lo[c]; o[9 + c] = b1.hi[c]; }
    memcpy(&o[12], &r0, 4);
    memcpy(&o[13], &r1, 4);
    *out = b0;
    hb_grow(*out, b1);
    return (int)me;
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
    struct NodesGuard {  // early error returns below must not leak the node array
        float4** p; bool armed = true;
        ~NodesGuard() { if (armed && *p) cudaFree(*p); }
    } guard{&nodes};

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

    // PLOC scratch
    int *ref[2] = {nullptr, nullptr}, *nn = nullptr, *keep = nullptr, *pos = nullptr, *ctl = nullptr;
    float4 *lo[2] = {nullptr, nullptr}, *hi[2] = {nullptr, nullptr};
    void* scan_tmp = nullptr;
    size_t scan_bytes = 0;
    double* d_cost = nullptr;
    PT_CUDA(sc.alloc(&d_cost, 2));
    int* d_depth = nullptr;  // [slot] depth of the candidate built into that slot, [2] scratch of k_dfs_index
    PT_CUDA(sc.alloc(&d_depth, 3));
    PT_CUDA(cudaMemsetAsync(d_depth, 0, 3 * sizeof(int), st));

    // one hierarchy (Karras or PLOC) + refit into `out`; its SAH cost (sum of all child-box areas) into d_cost[slot]
    int *subtree = nullptr, *newid = nullptr;
    float4* nodes_tmp = nullptr;
    const char* denv = getenv("PT_PLOC_DFS");  // "0": keep PLOC's creation-order numbering (A/B runs)
    const bool dfs = !(denv && denv[0] == '0');
    const char* renv = getenv("PT_PLOC_RADIUS");  // search radius in Morton order (A/B runs)
    const int radius = renv && atoi(renv) > 0 ? atoi(renv) : PT_PLOC_RADIUS;
    auto build = [&](bool ploc, float4* out, int slot, bool index_keys = false) -> int {
        PT_CUDA(cudaMemsetAsync(flags, 0, (size_t)(n - 1) * sizeof(int), st));
        PT_CUDA(cudaMemsetAsync(d_depth + slot, 0, sizeof(int), st));
        PT_CUDA(cudaMemsetAsync(d_cost + slot, 0, sizeof(double), st));
        if (ploc) {
            if (!nn) {
                for (int b = 0; b < 2; ++b) {
                    PT_CUDA(sc.alloc(&ref[b], (size_t)n));
                    PT_CUDA(sc.alloc(&lo[b], (size_t)n));
                    PT_CUDA(sc.alloc(&hi[b], (size_t)n));
                }
                PT_CUDA(sc.alloc(&nn, (size_t)n));
                PT_CUDA(sc.alloc(&keep, (size_t)n));
                PT_CUDA(sc.alloc(&pos, (size_t)n));
                PT_CUDA(sc.alloc(&ctl, 2));  // [0] next node index, [1] cluster count after the round
                PT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, pos, (int)n, st));
                PT_CUDA(sc.alloc((char**)&scan_tmp, scan_bytes));
            }
            const int ctl0[2] = {(int)(n - 2), (int)n};
            PT_CUDA(cudaMemcpyAsync(ctl, ctl0, sizeof ctl0, cudaMemcpyHostToDevice, st));
            k_ploc_init<<<gridN, B, 0, st>>>(n, vb.Current(), d_prim_aabb, ref[0], lo[0], hi[0]);
            int c = (int)n, cur = 0, rounds = 0;
            while (c > 1) {
                const unsigned g = (unsigned)((c + B - 1) / B);
                k_ploc_nn<<<g, B, 0, st>>>(c, radius, lo[cur], hi[cur], nn);
                k_ploc_merge<<<g, B, 0, st>>>(c, n, nn, ref[cur], lo[cur], hi[cur], keep, ctl, children, parent);
                PT_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, keep, pos, c, st));
                k_ploc_compact<<<g, B, 0, st>>>(c, keep, pos, ref[cur], lo[cur], hi[cur], ref[cur ^ 1], lo[cur ^ 1], hi[cur ^ 1], ctl + 1);
                int c_new = 0;
                PT_CUDA(cudaMemcpyAsync(&c_new, ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
                PT_CUDA(cudaStreamSynchronize(st));
                if (c_new >= c || ++rounds > 4096) {  // chain-like inputs that merge one pair per round: not a candidate
                    pt_set_error("pt_lbvh_build: PLOC gave up (%d -> %d clusters after %d rounds)", c, c_new, rounds);
                    return PT_ERR_INVALID;
                }
                c = c_new;
                cur ^= 1;
            }
            const int minus1 = -1;
            PT_CUDA(cudaMemcpy(parent, &minus1, sizeof(int), cudaMemcpyHostToDevice));  // the root (node 0) has no parent
        } else {
            if (index_keys) k_index_keys<<<gridN, B, 0, st>>>(n, kb.Current());  // the Morton keys are not needed again
            k_hierarchy<<<(unsigned)((n - 1 + B - 1) / B), B, 0, st>>>(kb.Current(), n, children, parent);
        }
        k_tree_depth<<<gridN, B, 0, st>>>(n, parent, d_depth + slot);
        if (ploc && dfs) {  // PLOC numbers nodes by merge round: renumber depth-first (see k_dfs_index)
            if (!subtree) {
                PT_CUDA(sc.alloc(&subtree, (size_t)(n - 1)));
                PT_CUDA(sc.alloc(&newid, (size_t)(n - 1)));
                PT_CUDA(sc.alloc(&nodes_tmp, (size_t)(n - 1) * 4));
            }
            const unsigned gi = (unsigned)((n - 1 + B - 1) / B);
            k_refit_emit<<<gridN, B, 0, st>>>(n, children, parent, vb.Current(), d_prim_aabb, node_box, flags, nodes_tmp, subtree);
            k_dfs_index<<<gi, B, 0, st>>>(n - 1, children, parent, subtree, newid, d_depth + 2);
            k_relayout<<<gi, B, 0, st>>>(n - 1, nodes_tmp, newid, out);
        } else {
            k_refit_emit<<<gridN, B, 0, st>>>(n, children, parent, vb.Current(), d_prim_aabb, node_box, flags, out, nullptr);
        }
        k_sah_cost<<<(unsigned)((n - 1 + B - 1) / B), B, 0, st>>>(out, n - 1, d_cost + slot);
        PT_CUDA(cudaStreamSynchronize(st));
        PT_CUDA(cudaGetLastError());
        return PT_OK;
    };

    // Which hierarchy: PT_BUILDER=lbvh|ploc forces one (A/B runs).  Otherwise trees of up to PT_PLOC_MAX primitives are
    // built BOTH ways (a millisecond each at these sizes) and the one with the lower SAH cost is kept — PLOC wins on the
    // triangle meshes (Yoimiya 8.3 instead of 9.6 node visits per segment), the Karras tree on the RTIOW sphere field.
    const char* env = getenv("PT_BUILDER");
    const bool force = env && env[0];
    PT_CUDA(cudaMemsetAsync(d_cost, 0, 2 * sizeof(double), st));
    int rcb;
    // host full-sweep SAH (tiny trees): boxes down, nodes up, cost through the same kernel
    auto host_sah = [&](float4* out, double* cost_out) -> int {
        std::vector<int> prim((size_t)n);
        PT_CUDA(cudaMemcpy(prim.data(), d_local_ids, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<HostBox> box((size_t)n);
        {
            int max_p = 0;
            for (int p : prim) max_p = std::max(max_p, p);
            std::vector<float4> all((size_t)(max_p + 1) * 2);
            PT_CUDA(cudaMemcpy(all.data(), d_prim_aabb, all.size() * sizeof(float4), cudaMemcpyDeviceToHost));
            for (int64_t k = 0; k < n; ++k) {
                const float4 a = all[2 * (size_t)prim[k]], b = all[2 * (size_t)prim[k] + 1];
                box[k].lo[0] = a.x; box[k].lo[1] = a.y; box[k].lo[2] = a.z;
                box[k].hi[0] = b.x; box[k].hi[1] = b.y; box[k].hi[2] = b.z;
            }
        }
        std::vector<int> idx((size_t)n);
        for (int64_t k = 0; k < n; ++k) idx[k] = (int)k;
        std::vector<float> hn, tmp_area((size_t)n);
        hn.reserve((size_t)(n - 1) * 16);
        HostBox root_box;
        int depth = 0;
        host_sah_rec(idx, 0, (int)n, box, prim, hn, tmp_area, &root_box, 0, &depth);
        PT_REQUIRE((int64_t)hn.size() == (n - 1) * 16, "host SAH emitted a wrong node count");
        if (depth > PT_MAX_TREE_DEPTH) {  // would not fit the traversal stack: not a candidate
            *cost_out = INFINITY;
            return PT_OK;
        }
        PT_CUDA(cudaMemcpy(out, hn.data(), hn.size() * sizeof(float), cudaMemcpyHostToDevice));
        PT_CUDA(cudaMemsetAsync(d_cost + 1, 0, sizeof(double), st));
        k_sah_cost<<<(unsigned)((n - 1 + B - 1) / B), B, 0, st>>>(out, n - 1, d_cost + 1);
        PT_CUDA(cudaMemcpyAsync(cost_out, d_cost + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        PT_CUDA(cudaStreamSynchronize(st));
        return PT_OK;
    };
    // every candidate's depth is measured (k_tree_depth): one that does not fit the traversal stack is not a candidate
    auto depth_of = [&](int slot, int* out) -> int {
        PT_CUDA(cudaMemcpy(out, d_depth + slot, sizeof(int), cudaMemcpyDeviceToHost));
        return PT_OK;
    };
    bool have = false;  // `nodes` holds a tree of at most PT_MAX_TREE_DEPTH levels
    int depth = 0;
    if (force && env[0] == 's') {
        double c = 0.0;
        rcb = host_sah(nodes, &c);
        have = rcb == PT_OK && c < INFINITY;
    } else if (force || n > PT_PLOC_MAX) {
        rcb = build(force && env[0] == 'p', nodes, 0);
        if (rcb == PT_OK) rcb = depth_of(0, &depth);
        have = rcb == PT_OK && depth <= PT_MAX_TREE_DEPTH;
    } else {
        float4* nodes2 = nullptr;
        cudaError_t ea = cudaMalloc(&nodes2, (size_t)(n - 1) * 4 * sizeof(float4));
        PT_CUDA(ea);
        double cost[2] = {INFINITY, INFINITY};
        int dep[2] = {0, 0};
        rcb = build(false, nodes, 0);
        if (rcb == PT_OK) {
            const int rcp = build(true, nodes2, 1);  // PLOC is an optional candidate: its failure keeps the Karras tree
            if (cudaMemcpy(cost, d_cost, sizeof cost, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(dep, d_depth, sizeof dep, cudaMemcpyDeviceToHost) != cudaSuccess) rcb = PT_ERR_CUDA;
            if (rcp != PT_OK) {
                cost[1] = INFINITY;
                if (rcp == PT_ERR_CUDA) rcb = rcp;
            }
        }
        for (int k = 0; k < 2; ++k)
            if (dep[k] > PT_MAX_TREE_DEPTH) cost[k] = INFINITY;
        if (rcb == PT_OK && cost[1] < cost[0]) { std::swap(nodes, nodes2); cost[0] = cost[1]; }
        double cost_sah = -1.0;
        if (rcb == PT_OK && n <= PT_HOST_SAH_MAX) {  // tiny tree: a third candidate from the host (nodes2 is free again)
            rcb = host_sah(nodes2, &cost_sah);
            if (rcb == PT_OK && cost_sah < cost[0]) { std::swap(nodes, nodes2); cost[0] = cost_sah; }
        }
        have = rcb == PT_OK && cost[0] < INFINITY;
        if (getenv("PT_BUILD_VERBOSE"))
            fprintf(stderr, "[libb200pt] %lld prims: SAH cost kept %.6g (lbvh depth %d, ploc %.6g depth %d, host sweep %.6g)\n",
                    (long long)n, cost[0], dep[0], cost[1], dep[1], cost_sah);
        cudaFree(nodes2);
    }
    if (rcb == PT_OK && (!have || getenv("PT_FORCE_INDEX_TREE"))) {
        // degenerate geometry (long runs of equal Morton codes, chains): the radix tree over the sorted index always fits
        rcb = build(false, nodes, 0, true);
        if (rcb == PT_OK) rcb = depth_of(0, &depth);
        if (rcb == PT_OK && depth > PT_MAX_TREE_DEPTH) {
            pt_set_error("pt_lbvh_build: index tree is %d levels deep", depth);
            rcb = PT_ERR_INVALID;
        }
    }
    if (rcb) return rcb;  // the guard frees `nodes`
    guard.armed = false;
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
