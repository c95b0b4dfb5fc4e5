// pt_math.cuh — float3 helpers and the counter-based RNG shared by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PT_PI 3.14159265358979323846f
#define PT_EPS 1e-4f  // reference epsilon: 10_final/world.py:30, legacy 15_module.py:44

#define PT_DEV __device__ __forceinline__

PT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PT_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
PT_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PT_DEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
PT_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PT_DEV float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
PT_DEV float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PT_DEV float3 normalize(float3 a) { return a * rsqrtf(dot(a, a)); }
PT_DEV float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PT_DEV float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PT_DEV float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

// pcg4d (Jarzynski & Olano 2020): 4x32 -> 4x32 bits.  Key = (pixel, sample, stream, seed);
// stream 0 = camera ray, 1+2b / 2+2b = bounce b.  Uniform = top 24 bits * 2^-24 in [0,1).
// Integer arithmetic only, so the CPU oracle draws bit-identical uniforms.
PT_DEV uint4 pcg4d(uint4 v) {
    v.x = v.x * 1664525u + 1013904223u;
    v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u;
    v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
PT_DEV float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
PT_DEV float4 rng4(uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t seed) {
    uint4 v = pcg4d(make_uint4(pixel, sample, stream, seed));
    return make_float4(u01(v.x), u01(v.y), u01(v.z), u01(v.w));
}

// sin/cos of 2*pi*u, u in [0,1): shift into [-pi, pi) where the MUFU approximations are accurate to
// ~2^-21 absolute; sin(x + pi) = -sin x, cos(x + pi) = -cos x.
PT_DEV void sincos_2pi(float u, float* s, float* c) {
    float x = fmaf(2.0f * PT_PI, u, -PT_PI);
    float ss, cc;
    __sincosf(x, &ss, &cc);
    *s = -ss;
    *c = -cc;
}

// asin / atan2 of the environment lookup (15_module.py:970-977) as degree-13 / degree-7 minimax-style polynomials
// (least squares on Chebyshev nodes; maximum error 2.7e-7 rad and 1.8e-7 rad against double precision: 1e-4 texels of a
// 2048-texel environment).  About 20 and 14 instructions instead of the 40 and 25 of the CUDA math library versions.
PT_DEV float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    const float s = a * a;
    float r = fmaf(s, 0.0068426248617470264f, -0.03372593969106674f);
    r = fmaf(r, s, 0.0798112079501152f);
    r = fmaf(r, s, -0.13247522711753845f);
    r = fmaf(r, s, 0.19813214242458344f);
    r = fmaf(r, s, -0.33318302035331726f);
    r = fmaf(r, s, 0.9999966621398926f);
    r *= a;
    if (ay > ax) r = 0.5f * PT_PI - r;
    if (x < 0.0f) r = PT_PI - r;
    return copysignf(r, y);
}
PT_DEV float fast_asinf(float x) {
    const float a = fminf(fabsf(x), 1.0f);
    float q = fmaf(a, 0.002251368248835206f, -0.011012386530637741f);
    q = fmaf(q, a, 0.02674933150410652f);
    q = fmaf(q, a, -0.048724401742219925f);
    q = fmaf(q, a, 0.08873733133077621f);
    q = fmaf(q, a, -0.21458369493484497f);
    q = fmaf(q, a, 1.5707961320877075f);
    return copysignf(0.5f * PT_PI - sqrtf(1.0f - a) * q, x);
}
