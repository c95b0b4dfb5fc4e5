// wf_common.cuh — constants and POD parameter blocks shared by the render kernels (wavefront.cu: split and
// K-step fused wavefront; persist.cu: persistent while-while kernels).
#pragma once
#include "shade.cuh"

#define PT_BLOCK 256
#define PT_WARPS (PT_BLOCK / 32)

// device counter block (unsigned long long[16])
#define CNT_NEXT_PATH 0
#define CNT_SEGMENTS 1
#define CNT_NODES 2
#define CNT_PRIMS 3
#define CNT_QUEUE 8  // [8], [9]: ping-pong queue sizes (low 32 bits used)
#define CNT_WORDS 16
#define CNT_LIVE 16  // fused mode: live-path count after launch L at [CNT_LIVE + L]
#define PT_MAX_LAUNCHES 4080
#define CNT_TOTAL_WORDS (CNT_LIVE + PT_MAX_LAUNCHES)
#define PT_MODE_AUTO 0
#define PT_MODE_SPLIT 1
#define PT_MODE_FUSED 2
#define PT_MODE_PERSIST 3
#define PT_MODE_QUEUE 4
#define PT_MODE_DUAL 5

struct RenderConsts {
    CameraDev cam;
    unsigned long long total_paths;
    int W, H;
    uint32_t seed, spp_offset;
    int max_depth, shading_model;
    float absorptivity, tmin;
    unsigned pool_cap;
    int accum_sq;
    // fused mode: static striding of path ids over the pool, id -> (sample, pixel) without division
    uint32_t stride_samples, stride_pixels;  // pool_cap = stride_samples * W*H + stride_pixels
    uint32_t sample_end;                     // spp_offset + spp
    int row0, row1;                          // rows [row0, row1) of the image are rendered (persistent kernel; else 0, H)
    uint32_t unit_samples;                   // persistent kernel: samples per work unit (one 8x4 tile x unit_samples)
};

struct PoolPtrs {
    float4 *o, *d, *l;
};


// Camera.get_rays for one path (camera.py:71-93, 15_module.py:438-453): fused ray generation
PT_DEV void start_path(const RenderConsts& rc, PathState& p, uint32_t pixel, uint32_t sample) {
    p.pixel = pixel;
    p.sample = sample;
    p.bounce = 0u;
    p.l = f3(1.0f, 1.0f, 1.0f);
    const float4 u = rng4(pixel, sample, 0u, rc.seed);
    const uint32_t j = pixel / (uint32_t)rc.W;
    camera_ray(rc.cam, (int)(pixel - j * (uint32_t)rc.W), (int)j, u, &p.o, &p.d);
}
