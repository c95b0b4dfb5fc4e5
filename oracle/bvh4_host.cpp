// CPU harness around the product's host-side BVH2 -> BVH4 collapse (learn_path_tracing_b200/csrc/bvh4.h) so that
// tests/test_bvh4_host.py can check it without a GPU.  Test infrastructure, like everything under oracle/.
#include "../learn_path_tracing_b200/csrc/bvh4.h"

extern "C" long long bvh4_collapse_host(const float* nodes16, long long n_nodes, float* out, long long cap_nodes) {
    const std::vector<float> w = bvh4::collapse(nodes16, n_nodes);
    const long long m = (long long)(w.size() / BVH4_NODE_FLOATS);
    if (m > cap_nodes) return -1;
    memcpy(out, w.data(), w.size() * sizeof(float));
    return m;
}
