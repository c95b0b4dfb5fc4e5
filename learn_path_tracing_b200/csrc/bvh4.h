// bvh4.h — collapse of the emitted BVH2 (k_refit_emit layout) into a 4-wide tree.  EXPERIMENTAL, opt-in (PT_WIDE=1 in the
// environment at build time AND PT_FLAG_TRACE_WIDE on the trace call).  Validated on the CPU (tests/test_bvh4_host.py:
// the oracle walking the collapsed tree returns the BVH2's hits) and on the B200 with the last GPU seconds of round 1
// (tests/test_gpu_parity.py::test_wide_tree_traversal_...: hit records bit-identical to the binary walk, 40.5 instead of
// 71.4 steps per ray, 0.62 instead of 0.82 ms on 400 k rays / 200 k triangles, untuned).  Not yet measured at the
// benchmark sizes, not yet in the render kernels: first item of the next round.
//
// Why: tools/bvh4_proto.py on the Yoimiya PLOC tree — 23 424 binary nodes become 11 342 wide nodes of average arity 3.07,
// traversal steps fall to 0.51x (secondary rays) / 0.64x (primary), the total number of box tests falls too (55 vs 69
// per secondary ray) because intermediate boxes disappear, triangle tests stay.
//
// Collapse rule: the children of a wide node start as the two children of a binary node; while there are fewer than
// four, the INNER child with the largest surface area is replaced by its own two children.  Nodes are emitted
// depth-first (root = 0).
//
// Wide node = 128 bytes = four 256-bit loads:
//   float  minx[4], miny[4]  |  minz[4], maxx[4]  |  maxy[4], maxz[4]  |  int ref[4], pad[4]
//   ref >= 0 wide node index, ref < 0 leaf with primitive ~ref, ref == BVH4_EMPTY unused slot (box = 0: ignored by ref)
//
// Pure host C++ (no CUDA) so that the same header builds into the CPU test harness (oracle/Makefile: libbvh4host.so).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#define BVH4_EMPTY 0x7fffffff
#define BVH4_NODE_FLOATS 32

namespace bvh4 {

struct Child {
    float lo[3], hi[3];
    int ref;  // BVH2 encoding: >= 0 inner node, < 0 leaf ~prim
};

inline float area(const Child& c) {
    const float x = c.hi[0] - c.lo[0], y = c.hi[1] - c.lo[1], z = c.hi[2] - c.lo[2];
    return x * y + y * z + z * x;
}

inline void children_of(const float* nodes16, int i, Child out[2]) {
    const float* n = nodes16 + 16 * (size_t)i;
    for (int k = 0; k < 2; ++k) {
        for (int c = 0; c < 3; ++c) { out[k].lo[c] = n[6 * k + c]; out[k].hi[c] = n[6 * k + 3 + c]; }
        memcpy(&out[k].ref, &n[12 + k], 4);
    }
}

// nodes16: n_nodes x 16 floats (root = node 0).  Returns the wide nodes (BVH4_NODE_FLOATS floats each, root = 0).
inline std::vector<float> collapse(const float* nodes16, int64_t n_nodes) {
    std::vector<float> out;
    if (n_nodes <= 0) return out;
    out.reserve((size_t)n_nodes / 2 * BVH4_NODE_FLOATS + BVH4_NODE_FLOATS);
    struct Todo { int bvh2; size_t wide; int slot; };  // wide node `wide` slot `slot` waits for the index of bvh2's wide node
    std::vector<Todo> stack;
    stack.push_back({0, (size_t)-1, 0});
    while (!stack.empty()) {
        const Todo t = stack.back();
        stack.pop_back();
        const size_t me = out.size() / BVH4_NODE_FLOATS;
        if (t.wide != (size_t)-1) {
            const int r = (int)me;
            memcpy(&out[t.wide * BVH4_NODE_FLOATS + 24 + t.slot], &r, 4);
        }
        Child ch[4];
        int cnt = 2;
        children_of(nodes16, t.bvh2, ch);
        while (cnt < 4) {
            int best = -1;
            float best_a = -1.0f;
            for (int k = 0; k < cnt; ++k)
                if (ch[k].ref >= 0 && area(ch[k]) > best_a) { best_a = area(ch[k]); best = k; }
            if (best < 0) break;
            Child two[2];
            children_of(nodes16, ch[best].ref, two);
            ch[best] = two[0];
            ch[cnt++] = two[1];
        }
        out.resize(out.size() + BVH4_NODE_FLOATS, 0.0f);
        float* o = &out[me * BVH4_NODE_FLOATS];
        for (int k = 0; k < 4; ++k) {
            int ref = BVH4_EMPTY;
            if (k < cnt) {
                o[0 + k] = ch[k].lo[0]; o[4 + k] = ch[k].lo[1]; o[8 + k] = ch[k].lo[2];
                o[12 + k] = ch[k].hi[0]; o[16 + k] = ch[k].hi[1]; o[20 + k] = ch[k].hi[2];
                ref = ch[k].ref < 0 ? ch[k].ref : 0;  // inner refs are patched when the child is emitted
            }
            memcpy(&o[24 + k], &ref, 4);
        }
        // children are emitted depth-first: push in reverse so that slot 0's subtree follows its parent directly
        for (int k = cnt - 1; k >= 0; --k)
            if (ch[k].ref >= 0) stack.push_back({ch[k].ref, me, k});
    }
    return out;
}

}  // namespace bvh4
