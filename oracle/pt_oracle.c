/*
 * pt_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See pt_oracle.h.
 *
 * Plain C11 + OpenMP, float32 arithmetic, compiled with -ffp-contract=off so every expression is
 * evaluated exactly as written (no FMA contraction).  Every function cites the reference lines it
 * restates; "ref:" paths are relative to the reference checkout, with
 *   v2     = taichi_pathtracer/10_final
 *   legacy = legacy/PT_in_one_weekend/15_module.py
 *
 * What is NOT the reference's: the random numbers.  Taichi's ti.random is a per-thread xorshift
 * stream that nothing in the reference pins; the oracle (and the CUDA library) use a counter-based
 * generator keyed on (seed, pixel, sample, stream) so that CPU and GPU draw the same uniforms.
 */
#include "pt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846f
#define ORC_EPS 1e-4f /* ref: legacy:44 epsilon; v2 world.py:30 literal 1e-4 */

typedef struct { float x, y, z; } V3;

static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 vmul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 vscale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float vdot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 vcross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V3 vnormalized(V3 a) { /* Taichi Vector.normalized(): v / sqrt(v.dot(v)) */
    float n = sqrtf(vdot(a, a));
    return v3(a.x / n, a.y / n, a.z / n);
}
static inline V3 vload(const float* p) { return v3(p[0], p[1], p[2]); }

/* ------------------------------------------------------------------------------------------ */
/* counter-based RNG: pcg4d (Jarzynski & Olano 2020), 4 x 32 bit in -> 4 x 32 bit out;        */
/* uniform = top 24 bits * 2^-24 in [0,1).  Streams: 0 = camera ray; 1+2b, 2+2b = bounce b.    */
/* ------------------------------------------------------------------------------------------ */
static inline void pcg4d(uint32_t v[4]) {
    for (int i = 0; i < 4; ++i) v[i] = v[i] * 1664525u + 1013904223u;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
    for (int i = 0; i < 4; ++i) v[i] ^= v[i] >> 16;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}

void orc_rng4(uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t seed, float out[4]) {
    uint32_t v[4] = {pixel, sample, stream, seed};
    pcg4d(v);
    for (int i = 0; i < 4; ++i) out[i] = (float)(v[i] >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------------------------------ */
/* camera — ref: v2 camera.py:29-35 (sample_in_disk), :71-93 (get_rays); legacy:438-453        */
/* ------------------------------------------------------------------------------------------ */
static inline void camera_ray(const PtCamera* c, int W, int H, int i, int j, const float u[4], int grid, V3* ro, V3* rd) {
    V3 dir = vload(c->front), wa = vload(c->right), ha = vload(c->up), pos = vload(c->pos);
    float fx = ((float)i + u[0]) / (float)W - 0.5f; /* camera.py:88 */
    float fy = ((float)j + u[1]) / (float)H - 0.5f;
    if (grid) { /* stages 2-4, 2_camera_and_ray/camera.py:67: i / (width - 1) - 0.5, no jitter */
        fx = (float)i / (float)(W - 1) - 0.5f;
        fy = (float)j / (float)(H - 1) - 0.5f;
    }
    V3 target = vscale(vadd(vadd(dir, vscale(wa, fx * c->view_w)), vscale(ha, fy * c->view_h)), c->focal_length);
    float r = sqrtf(u[2]); /* camera.py:31-34 */
    float theta = 2.0f * ORC_PI * u[3];
    float sx = r * cosf(theta), sy = r * sinf(theta);
    V3 origin = vscale(vadd(vscale(wa, sx), vscale(ha, sy)), c->aperture / 2.0f); /* camera.py:90 */
    *ro = vadd(pos, origin);
    *rd = vnormalized(vsub(target, origin));
}

void orc_generate_rays(const PtCamera* cam, int width, int height, int sample, uint32_t seed, float* rays) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < height; ++j)
        for (int i = 0; i < width; ++i) {
            uint32_t pix = (uint32_t)j * (uint32_t)width + (uint32_t)i;
            float u[4];
            orc_rng4(pix, (uint32_t)sample, 0u, seed, u);
            V3 ro, rd;
            camera_ray(cam, width, height, i, j, u, 0, &ro, &rd);
            float* r = rays + (size_t)pix * 8;
            r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = ORC_EPS;
            r[4] = rd.x; r[5] = rd.y; r[6] = rd.z; r[7] = INFINITY;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* v2 intersection — ref: world.py:43-60 (Sphere.hit), :24-34 (World.hit)                      */
/* ------------------------------------------------------------------------------------------ */
/* returns t (or -1); literal restatement incl. the "near root < 1e-4 and transparent -> far root" rule */
static inline float sphere_hit_t(V3 ro, V3 rd, V3 center, float radius, int transparency) {
    V3 oc = vsub(ro, center);
    float b = 2.0f * vdot(oc, rd);
    float c = vdot(oc, oc) - radius * radius;
    float disc = b * b - 4.0f * c;
    float t = -1.0f;
    if (disc >= 0.0f) {
        float sq = sqrtf(disc);
        t = (-b - sq) / 2.0f;
        if (t < ORC_EPS && transparency) t = (-b + sq) / 2.0f;
    }
    return t;
}
static inline double sphere_hit_t64(const float* ro, const float* rd, const float* cr, int transparency) {
    double ocx = (double)ro[0] - cr[0], ocy = (double)ro[1] - cr[1], ocz = (double)ro[2] - cr[2];
    double b = 2.0 * (ocx * rd[0] + ocy * rd[1] + ocz * rd[2]);
    double c = ocx * ocx + ocy * ocy + ocz * ocz - (double)cr[3] * cr[3];
    double disc = b * b - 4.0 * c;
    double t = -1.0;
    if (disc >= 0.0) {
        double sq = sqrt(disc);
        t = (-b - sq) / 2.0;
        if (t < 1e-4 && transparency) t = (-b + sq) / 2.0;
    }
    return t;
}

/* closest sphere: first wins ties (strict <), accepted iff t >= 1e-4 (world.py:30) */
static inline int world_hit_v2(const float* cr, const PtMaterial* mats, int n, V3 ro, V3 rd, float* t_out) {
    int best = -1;
    float bt = -1.0f;
    for (int i = 0; i < n; ++i) {
        float t = sphere_hit_t(ro, rd, vload(cr + 4 * i), cr[4 * i + 3], mats[i].transparency);
        if (t >= ORC_EPS && (bt < 0.0f || t < bt)) { bt = t; best = i; }
    }
    *t_out = bt;
    return best;
}

void orc_trace_spheres(const float* cr, const PtMaterial* mats, int n, const float* rays, int64_t nrays,
                       int32_t* prim_id, float* t, double* t64) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nrays; ++k) {
        const float* r = rays + 8 * k;
        float bt;
        int id = world_hit_v2(cr, mats, n, vload(r), vload(r + 4), &bt);
        prim_id[k] = id;
        t[k] = bt;
        if (t64) t64[k] = id >= 0 ? sphere_hit_t64(r, r + 4, cr + 4 * id, mats[id].transparency) : -1.0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* v2 BSDFs — ref: bsdf.py:5-110                                                               */
/* ------------------------------------------------------------------------------------------ */
static inline V3 sample_at_sphere(float u0, float u1) { /* bsdf.py:5-12; legacy:295-302 */
    float z = 1.0f - 2.0f * u0;
    float r = sqrtf(1.0f - z * z);
    float theta = 2.0f * ORC_PI * u1;
    return v3(r * cosf(theta), r * sinf(theta), z);
}
static inline V3 sample_lambertian(V3 n, float u0, float u1) { /* bsdf.py:15-18 */
    return vnormalized(vadd(n, sample_at_sphere(u0, u1)));
}
static inline V3 slerp(V3 a, V3 b, float t) { /* bsdf.py:21-27 */
    float d = vdot(a, b);
    d = fminf(fmaxf(d, -1.0f), 1.0f);
    float omega = acosf(d);
    float so = sinf(omega);
    V3 o;
    if (so < 1e-6f) o = vadd(vscale(a, 1.0f - t), vscale(b, t));
    else o = vadd(vscale(a, sinf((1.0f - t) * omega) / so), vscale(b, sinf(t * omega) / so));
    return vnormalized(o);
}
static inline V3 reflect_v2(V3 dir, V3 n) { /* bsdf.py:40-44 */
    float k = -vdot(dir, n);
    return vadd(dir, vscale(n, 2.0f * k));
}
static inline V3 sample_normal(V3 dir, V3 normal, float roughness, float u0, float u1) { /* bsdf.py:30-37 */
    V3 s = sample_lambertian(normal, u0, u1);
    V3 r = reflect_v2(dir, normal);
    r = slerp(r, s, roughness * roughness);
    return vnormalized(vsub(r, dir));
}
static inline V3 refract_v2(V3 dir, V3 n, float ior) { /* bsdf.py:47-59 */
    float k = vdot(dir, n);
    V3 perp = v3((dir.x - k * n.x) / ior, (dir.y - k * n.y) / ior, (dir.z - k * n.z) / ior);
    float len2 = vdot(perp, perp);
    if (len2 > 1.0f) return reflect_v2(dir, n);
    float kk = sqrtf(1.0f - len2);
    return vadd(perp, vscale(n, -kk));
}
static inline float pow5(float x) { return powf(x, 5.0f); }

/* one scatter event of stages 7-10; returns 0 (always continues).  u = stream 1+2b, u2 = stream 2+2b */
static inline void scatter_v2(const PtMaterial* m, float ior, V3 point, V3 normal, V3* ro, V3* rd, V3* l,
                              const float u[4], const float u2[4]) {
    V3 d = *rd;
    V3 n = sample_normal(d, normal, m->roughness, u[0], u[1]);
    float cos_theta = fmaxf(0.0f, vdot(n, vneg(d)));
    if (m->metallic == 1) { /* __main__.py:70; MetalBSDF bsdf.py:71-86 */
        V3 F0 = vload(m->albedo);
        float w = pow5(1.0f - cos_theta);
        V3 F = v3(F0.x + (1.0f - F0.x) * w, F0.y + (1.0f - F0.y) * w, F0.z + (1.0f - F0.z) * w);
        *l = vmul(*l, F);
        *ro = point;
        *rd = reflect_v2(d, n);
    } else { /* DielectricBSDF bsdf.py:89-110 */
        float F0 = ((ior - 1.0f) / (ior + 1.0f)) * ((ior - 1.0f) / (ior + 1.0f));
        float F = F0 + (1.0f - F0) * pow5(1.0f - cos_theta);
        *ro = point;
        if (u[2] > F) {
            *l = vmul(*l, vload(m->albedo));
            if (m->transparency) *rd = refract_v2(d, n, ior);
            else *rd = sample_lambertian(normal, u2[0], u2[1]);
        } else {
            *rd = reflect_v2(d, n);
        }
    }
}

static inline V3 background_color(V3 rd) { /* __main__.py:58-62 */
    float t = 0.5f * (rd.y + 1.0f);
    return v3((1.0f - t) * 1.0f + t * 0.5f, (1.0f - t) * 1.0f + t * 0.7f, (1.0f - t) * 1.0f + t * 1.0f);
}

/* ------------------------------------------------------------------------------------------ */
/* legacy texture sampling — ref: legacy:238-258 (bilinear), :65-115 (load_texture transfer fns) */
/* ------------------------------------------------------------------------------------------ */
typedef struct { V3 albedo; V3 normal; float roughness; float metallic; } Texel;

static float g_lut_albedo[256], g_lut_sq[256], g_lut_nrm[256];
static int g_lut_ready = 0;
static void init_luts(void) {
    if (g_lut_ready) return;
    for (int i = 0; i < 256; ++i) {
        double x = i / 255.0;
        g_lut_albedo[i] = (float)pow(x, 2.2); /* legacy:101 albedo**2.2 */
        g_lut_sq[i] = (float)(x * x);         /* legacy:102-103 roughness**2, metallic**2 */
        g_lut_nrm[i] = (float)(x * 2.0 - 1.0); /* legacy:104 */
    }
    g_lut_ready = 1;
}

static inline int pymod(int a, int m) { /* ti.mod: Python modulo, result has the divisor's sign */
    int r = a % m;
    return (r != 0 && ((r < 0) != (m < 0))) ? r + m : r;
}

static inline Texel fetch_texel(const OrcScene* sc, int x, int y) {
    Texel t;
    memset(&t, 0, sizeof t);
    if (x < 0 || y < 0 || x >= sc->tex_W || y >= sc->tex_H) return t; /* outside the field: zero */
    const uint8_t* p = sc->texels + ((size_t)x * sc->tex_H + y) * 8;
    t.albedo = v3(g_lut_albedo[p[0]], g_lut_albedo[p[1]], g_lut_albedo[p[2]]);
    t.roughness = g_lut_sq[p[3]];
    t.normal = v3(g_lut_nrm[p[4]], g_lut_nrm[p[5]], g_lut_nrm[p[6]]);
    t.metallic = g_lut_sq[p[7]];
    return t;
}

typedef struct { int l, r, b, t; float lb, lt, rb, rt; } BilinearTaps;

/* legacy:238-258 incl. its quirks: truncating cast after -0.5, and v wrapped with the area WIDTH */
static inline BilinearTaps bilinear_taps(const int32_t* area, float u, float v) {
    int w = area[2] - area[0];
    int h = area[3] - area[1];
    BilinearTaps k;
    u = u * (float)w;
    v = v * (float)h;
    u = u - 0.5f;
    v = v - 0.5f;
    int l = (int)u, r = l + 1, b = (int)v, t = b + 1;
    k.lb = ((float)r - u) * ((float)t - v);
    k.lt = ((float)r - u) * (v - (float)b);
    k.rb = (u - (float)l) * ((float)t - v);
    k.rt = (u - (float)l) * (v - (float)b);
    k.l = area[0] + pymod(l, w);
    k.r = area[0] + pymod(r, w);
    k.b = area[1] + pymod(b, w); /* sic: width */
    k.t = area[1] + pymod(t, w);
    return k;
}

static inline Texel bilinear_texture(const OrcScene* sc, int id, float u, float v) {
    BilinearTaps k = bilinear_taps(sc->tex_areas + 4 * id, u, v);
    Texel a = fetch_texel(sc, k.l, k.b), b = fetch_texel(sc, k.l, k.t), c = fetch_texel(sc, k.r, k.b),
          d = fetch_texel(sc, k.r, k.t);
    Texel o;
    o.albedo = vadd(vadd(vadd(vscale(a.albedo, k.lb), vscale(b.albedo, k.lt)), vscale(c.albedo, k.rb)),
                    vscale(d.albedo, k.rt));
    o.normal = vadd(vadd(vadd(vscale(a.normal, k.lb), vscale(b.normal, k.lt)), vscale(c.normal, k.rb)),
                    vscale(d.normal, k.rt));
    o.roughness = ((k.lb * a.roughness + k.lt * b.roughness) + k.rb * c.roughness) + k.rt * d.roughness;
    o.metallic = ((k.lb * a.metallic + k.lt * b.metallic) + k.rb * c.metallic) + k.rt * d.metallic;
    if (sc->tex_flags && (sc->tex_flags[id] & 1)) o.normal = v3(0.0f, 0.0f, 1.0f); /* legacy:84,104 */
    return o;
}

static inline V3 fetch_env(const OrcScene* sc, int x, int y) {
    if (x < 0 || y < 0 || x >= sc->env_W || y >= sc->env_H) return v3(0, 0, 0);
    const float* p = sc->env + ((size_t)x * sc->env_H + y) * 3;
    return v3(p[0], p[1], p[2]);
}

static inline V3 environment_color(const OrcScene* sc, V3 rd) { /* legacy:970-977 */
    float phi = asinf(rd.y);
    float theta = atan2f(-rd.x, -rd.z);
    float u = (theta / ORC_PI + 1.0f) / 2.0f;
    float v = phi / ORC_PI + 0.5f;
    BilinearTaps k = bilinear_taps(sc->env_area, u, v);
    V3 a = fetch_env(sc, k.l, k.b), b = fetch_env(sc, k.l, k.t), c = fetch_env(sc, k.r, k.b), d = fetch_env(sc, k.r, k.t);
    return vadd(vadd(vadd(vscale(a, k.lb), vscale(b, k.lt)), vscale(c, k.rb)), vscale(d, k.rt));
}

/* ------------------------------------------------------------------------------------------ */
/* legacy intersection — ref: legacy:851-861 (aabb_hit), :864-896 (sphere_hit), :909-953 (triangle_hit) */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float t;
    int prim;   /* global primitive id */
    V3 point, normal;
    V3 albedo;
    float roughness, metallic, ior, absorptivity;
    int transparency;
} LegacyHit;

static inline int aabb_hit(const float* low, const float* high, V3 ro, V3 rd) { /* legacy:851-861 */
    V3 inv = v3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    V3 i = vmul(vsub(vload(low), ro), inv);
    V3 o = vmul(vsub(vload(high), ro), inv);
    float tmaxx = fmaxf(i.x, o.x), tmaxy = fmaxf(i.y, o.y), tmaxz = fmaxf(i.z, o.z);
    float tminx = fminf(i.x, o.x), tminy = fminf(i.y, o.y), tminz = fminf(i.z, o.z);
    float t1 = fminf(tmaxx, fminf(tmaxy, tmaxz));
    float t0 = fmaxf(tminx, fmaxf(tminy, tminz));
    return t1 > t0 - ORC_EPS && t1 > 0.0f;
}

/* geometric part of triangle_hit (legacy:909-928): returns t or -1, barycentrics in w[3] */
static inline float triangle_hit_t(V3 p1, V3 p2, V3 p3, V3 o, V3 d, float w[3]) {
    V3 N = vnormalized(vcross(vsub(p2, p1), vsub(p3, p1)));
    float t = (vdot(N, p1) - vdot(o, N)) / vdot(d, N);
    w[0] = w[1] = w[2] = -1.0f;
    if (t > ORC_EPS) {
        V3 P = vadd(o, vscale(d, t));
        float w1 = vdot(vcross(vsub(p3, p2), vsub(P, p2)), N) / vdot(vcross(vsub(p3, p2), vsub(p1, p2)), N);
        float w2 = vdot(vcross(vsub(p1, p3), vsub(P, p3)), N) / vdot(vcross(vsub(p1, p3), vsub(p2, p3)), N);
        float w3 = 1.0f - w1 - w2;
        w[0] = w1; w[1] = w2; w[2] = w3;
        if (w1 > 0.0f && w2 > 0.0f && w3 > 0.0f) return t;
    }
    return -1.0f;
}

/* full triangle_hit (legacy:909-953) for face f of mesh m */
static inline void triangle_hit_full(const OrcScene* sc, const OrcMesh* m, int f, V3 o, V3 d, float absorptivity,
                                     LegacyHit* rec) {
    const int32_t* F = m->faces + 10 * f;
    V3 p1 = vload(m->pos + 3 * F[0]), p2 = vload(m->pos + 3 * F[3]), p3 = vload(m->pos + 3 * F[6]);
    float w[3];
    rec->t = -1.0f;
    float t = triangle_hit_t(p1, p2, p3, o, d, w);
    if (t < 0.0f) return;
    rec->t = t;
    rec->point = vadd(o, vscale(d, t));
    V3 n1 = vload(m->nrm + 3 * F[1]), n2 = vload(m->nrm + 3 * F[4]), n3 = vload(m->nrm + 3 * F[7]);
    rec->normal = vnormalized(vadd(vadd(vscale(n1, w[0]), vscale(n2, w[1])), vscale(n3, w[2]))); /* :936 */
    const float *t1 = m->uv + 2 * F[2], *t2 = m->uv + 2 * F[5], *t3 = m->uv + 2 * F[8];
    float u = (w[0] * t1[0] + w[1] * t2[0]) + w[2] * t3[0]; /* :941-942 */
    float v = (w[0] * t1[1] + w[1] * t2[1]) + w[2] * t3[1];
    Texel tx = bilinear_texture(sc, F[9], u, v); /* :943 */
    rec->albedo = tx.albedo;                     /* normal map unused for triangles (:945) */
    rec->roughness = tx.roughness;
    rec->metallic = tx.metallic;
    rec->ior = 1.5f;
    rec->absorptivity = absorptivity;
    rec->transparency = 0;
}

/* legacy sphere_hit (legacy:864-896) */
static inline void tsphere_hit_full(const OrcScene* sc, int s, V3 ro, V3 rd, float absorptivity, LegacyHit* rec) {
    const float* cr = sc->tsph_cr + 4 * s;
    int transparency = sc->tsph_transparency[s];
    rec->t = sphere_hit_t(ro, rd, vload(cr), cr[3], transparency);
    if (rec->t == -1.0f) return; /* discriminant < 0 */
    rec->point = vadd(ro, vscale(rd, rec->t));
    V3 N = vnormalized(vsub(rec->point, vload(cr)));
    float r = sqrtf(N.x * N.x + N.z * N.z);
    V3 T = v3(N.z / r, 0.0f, -N.x / r);
    V3 B = v3(N.x * N.y, -r, N.z * N.y);
    float phi = asinf(N.y);
    float theta = atan2f(-N.x, -N.z);
    float u = (theta / ORC_PI + 1.0f) / 2.0f;
    float v = phi / ORC_PI + 0.5f;
    Texel tx = bilinear_texture(sc, sc->tsph_tex[s], 2.0f * u, 1.0f * v);
    V3 nc = tx.normal;
    rec->normal = vnormalized(vadd(vadd(vscale(T, nc.x), vscale(B, nc.y)), vscale(N, nc.z)));
    rec->albedo = tx.albedo;
    rec->roughness = tx.roughness;
    rec->metallic = tx.metallic;
    rec->ior = 1.5f;
    rec->absorptivity = absorptivity;
    rec->transparency = transparency;
}

/* MeshBVHTree.hit (legacy:756-779): unordered, unpruned explicit-stack traversal of the stored tree;
 * right child is popped first.  Falls back to a brute-force face loop when no tree is stored. */
static void mesh_hit(const OrcScene* sc, const OrcMesh* m, int prim_base, V3 ro, V3 rd, float absorptivity,
                     LegacyHit* res, uint64_t* counts) {
    res->t = -1.0f;
    res->prim = -1;
    LegacyHit rec;
    if (m->n_nodes <= 0) {
        for (int f = 0; f < m->nf; ++f) {
            triangle_hit_full(sc, m, f, ro, rd, absorptivity, &rec);
            if (rec.t > ORC_EPS && (res->t < 0.0f || rec.t < res->t)) { *res = rec; res->prim = prim_base + f; }
        }
        return;
    }
    int stack[64];
    int sp = 0;
    stack[0] = 0;
    while (sp >= 0) {
        int cur = stack[sp];
        if (counts) counts[0]++;
        if (aabb_hit(m->node_low + 3 * cur, m->node_high + 3 * cur, ro, rd)) {
            int data = m->node_data[cur];
            if (data >= 0) {
                LegacyHit leaf;
                leaf.t = -1.0f;
                leaf.prim = -1;
                for (int f = m->leaf_cut[data]; f < m->leaf_cut[data + 1]; ++f) { /* triangle_list_hit :956-967 */
                    if (counts) counts[1]++;
                    triangle_hit_full(sc, m, f, ro, rd, absorptivity, &rec);
                    if (rec.t > ORC_EPS && (leaf.t < 0.0f || rec.t < leaf.t)) { leaf = rec; leaf.prim = prim_base + f; }
                }
                if (leaf.t > ORC_EPS && (res->t < 0.0f || leaf.t < res->t)) *res = leaf;
                sp -= 1;
            } else {
                stack[sp] = m->node_left[cur];
                sp += 1;
                stack[sp] = m->node_right[cur];
            }
        } else {
            sp -= 1;
        }
    }
}

/* World.hit (legacy:838-848): sphere BVH result first, then every mesh, strict < */
static void world_hit_legacy(const OrcScene* sc, V3 ro, V3 rd, float absorptivity, LegacyHit* res, uint64_t* counts) {
    res->t = -1.0f;
    res->prim = -1;
    LegacyHit rec;
    /* SphereBVHTree.hit (legacy:636-656) returns the closest sphere with t > eps; a brute-force loop
     * in leaf order is result-equivalent up to exact-t ties */
    for (int s = 0; s < sc->n_tsph; ++s) {
        tsphere_hit_full(sc, s, ro, rd, absorptivity, &rec);
        if (rec.t > ORC_EPS && (res->t < 0.0f || rec.t < res->t)) { *res = rec; res->prim = s; }
    }
    int base = sc->n_tsph;
    for (int k = 0; k < sc->n_mesh; ++k) {
        mesh_hit(sc, sc->meshes + k, base, ro, rd, absorptivity, &rec, counts);
        if (rec.t > ORC_EPS && (res->t < 0.0f || rec.t < res->t)) *res = rec;
        base += sc->meshes[k].nf;
    }
}

void orc_trace_legacy(const OrcScene* sc, const float* rays, int64_t nrays, int32_t* prim_id, float* t) {
    init_luts();
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < nrays; ++k) {
        LegacyHit h;
        world_hit_legacy(sc, vload(rays + 8 * k), vload(rays + 8 * k + 4), 0.25f, &h, NULL);
        prim_id[k] = h.t >= 0.0f ? h.prim : -1;
        t[k] = h.t >= 0.0f ? h.t : -1.0f;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* legacy shading — ref: legacy:281-347, :994-1013                                             */
/* ------------------------------------------------------------------------------------------ */
static inline V3 sample_in_sphere(float u0, float u1, float u2) { /* legacy:304-312 */
    float r = powf(u0, 1.0f / 3.0f);
    float theta = 2.0f * ORC_PI * u1;
    float phi = acosf(u2 * 2.0f - 1.0f);
    return v3(r * cosf(theta) * sinf(phi), r * sinf(theta) * sinf(phi), r * cosf(phi));
}
static inline V3 sample_reflect(V3 dir, V3 n, float roughness, V3 s) { /* legacy:329-334 */
    float k = -vdot(dir, n);
    V3 nd = vadd(dir, vscale(n, 2.0f * k));
    return vnormalized(vadd(nd, vscale(s, roughness)));
}
static inline V3 sample_refract(V3 dir, V3 n, float ior, float roughness, V3 s) { /* legacy:337-347 */
    float k = vdot(dir, n);
    V3 perp = v3((dir.x - k * n.x) / ior, (dir.y - k * n.y) / ior, (dir.z - k * n.z) / ior);
    float len2 = vdot(perp, perp);
    if (len2 > 1.0f) len2 = 1.0f;
    float kk = sqrtf(1.0f - len2);
    V3 nd = vadd(perp, vscale(n, -kk));
    return vnormalized(vadd(nd, vscale(s, roughness)));
}

/* gen_secondary_rays (legacy:994-1013). u = stream 1+2b (coin, coin, s0, s1), u2 = stream 2+2b (s2) */
static inline void scatter_legacy(const LegacyHit* h, V3* ro, V3* rd, V3* l, const float u[4], const float u2[4]) {
    V3 d = *rd, n = h->normal;
    if (u[0] < h->metallic) {
        float w = pow5(1.0f + vdot(n, d)); /* cal_reflectivity_metal :281-285 */
        V3 F0 = h->albedo;
        V3 F = v3(F0.x + (1.0f - F0.x) * w, F0.y + (1.0f - F0.y) * w, F0.z + (1.0f - F0.z) * w);
        *rd = sample_reflect(d, n, h->roughness, sample_in_sphere(u[2], u[3], u2[0]));
        *l = vmul(*l, F);
    } else {
        float f0 = ((h->ior - 1.0f) / (h->ior + 1.0f)) * ((h->ior - 1.0f) / (h->ior + 1.0f)); /* :288-292 */
        float F = f0 + (1.0f - f0) * pow5(1.0f + vdot(n, d));
        if (u[1] > F) {
            V3 k = vscale(h->albedo, 1.0f - h->absorptivity);
            if (h->transparency) *rd = sample_refract(d, n, h->ior, h->roughness, sample_in_sphere(u[2], u[3], u2[0]));
            else *rd = vnormalized(vadd(n, sample_at_sphere(u[2], u[3]))); /* sample_diffuse :322-325 */
            *l = vmul(*l, k);
        } else {
            *rd = sample_reflect(d, n, h->roughness, sample_in_sphere(u[2], u[3], u2[0]));
        }
    }
    *ro = vadd(h->point, vscale(n, 2.0f * ORC_EPS)); /* :1013 */
}

/* ------------------------------------------------------------------------------------------ */
/* legacy tutorial stages 6 / 7 — ref: legacy/PT_in_one_weekend/7_reflect.py:49-96,139-209,    */
/* 6_diffuse.py:106-170.  Same cal_reflectivity_*, sample_in_sphere, sample_diffuse as          */
/* 15_module.py:281-334; pinned on the reference's own 6_diffuse.png / 7_reflect.png            */
/* ------------------------------------------------------------------------------------------ */
/* World.hit (7_reflect.py:139-150): near root only (:160-176), accepted iff t > 1e-3, first wins ties */
static inline int world_hit_stage(const float* cr, int n, V3 ro, V3 rd, float* t_out) {
    int best = -1;
    float bt = -1.0f;
    for (int i = 0; i < n; ++i) {
        float t = sphere_hit_t(ro, rd, vload(cr + 4 * i), cr[4 * i + 3], 0);
        if (t > 1e-3f && (bt < 0.0f || t < bt)) { bt = t; best = i; }
    }
    *t_out = bt;
    return best;
}
static inline V3 sample_reflect_stage7(V3 dir, V3 n, float roughness, V3 s) { /* 7_reflect.py:91-96 */
    float k = -vdot(dir, n);
    V3 nd = vadd(dir, vscale(n, 2.0f * k));
    return vnormalized(vadd(nd, vscale(s, k * roughness)));
}
/* propagate_once hit branch: 7_reflect.py:187-204 (model STAGE7), 6_diffuse.py:164-167 (model STAGE6) */
static inline void scatter_stage(const PtMaterial* m, int model, float absorptivity, V3 point, V3 n, V3* ro, V3* rd, V3* l,
                                 const float u[4], const float u2[4]) {
    V3 d = *rd, albedo = vload(m->albedo);
    *ro = point;
    if (model == PT_SHADE_LEGACY_STAGE6) {
        *rd = vnormalized(vadd(n, sample_at_sphere(u[2], u[3])));
        *l = vmul(vscale(*l, absorptivity), albedo);
        return;
    }
    float w = pow5(1.0f + vdot(n, d));
    if (m->metallic) {
        V3 F = v3(albedo.x + (1.0f - albedo.x) * w, albedo.y + (1.0f - albedo.y) * w, albedo.z + (1.0f - albedo.z) * w);
        *rd = sample_reflect_stage7(d, n, m->roughness, sample_in_sphere(u[2], u[3], u2[0]));
        *l = vmul(*l, F);
    } else {
        float f0 = ((m->ior - 1.0f) / (m->ior + 1.0f)) * ((m->ior - 1.0f) / (m->ior + 1.0f));
        float F = f0 + (1.0f - f0) * w;
        if (u[1] > F) {
            *rd = vnormalized(vadd(n, sample_at_sphere(u[2], u[3])));
            *l = vmul(*l, vscale(albedo, absorptivity));
        } else {
            *rd = sample_reflect_stage7(d, n, m->roughness, sample_in_sphere(u[2], u[3], u2[0]));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* render — ref: v2 __main__.py:65-87,99-103; legacy:980-1036                                  */
/* ------------------------------------------------------------------------------------------ */
static inline int finite3(V3 c) { return isfinite(c.x) && isfinite(c.y) && isfinite(c.z); }

int orc_render(const OrcScene* sc, const PtCamera* cam, const PtRenderParams* p, float* accum, float* accum_sq,
               PtStats* stats, int threads) {
    const int W = p->width, H = p->height;
    if (W <= 0 || H <= 0 || p->spp < 0 || p->max_depth <= 0) return PT_ERR_INVALID;
    if ((p->flags & PT_FLAG_PIXEL_GRID) && (W < 2 || H < 2)) return PT_ERR_INVALID;
    init_luts();
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    else omp_set_num_threads(omp_get_num_procs());
#endif
    uint64_t tot_seg = 0, tot_nodes = 0, tot_prims = 0;
    const int count = (p->flags & PT_FLAG_COUNTERS) != 0;
    /* one pass over the image per sample, as the reference's host loop (__main__.py:99-103) */
    for (int s = 0; s < p->spp; ++s) {
        const uint32_t sample = (uint32_t)(p->spp_offset + s);
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot_seg, tot_nodes, tot_prims)
        for (int j = 0; j < H; ++j) {
            for (int i = 0; i < W; ++i) {
                const uint32_t pix = (uint32_t)j * (uint32_t)W + (uint32_t)i;
                float u[4], u2[4];
                orc_rng4(pix, sample, 0u, p->seed, u);
                V3 ro, rd, l = v3(1.0f, 1.0f, 1.0f);
                camera_ray(cam, W, H, i, j, u, (p->flags & PT_FLAG_PIXEL_GRID) != 0, &ro, &rd);
                int ended = 0;
                V3 radiance = v3(0, 0, 0);
                for (int b = 0; b < p->max_depth; ++b) {
                    tot_seg++;
                    orc_rng4(pix, sample, 1u + 2u * (uint32_t)b, p->seed, u);
                    orc_rng4(pix, sample, 2u + 2u * (uint32_t)b, p->seed, u2);
                    if (p->shading_model == PT_SHADE_LEGACY) {
                        LegacyHit h;
                        uint64_t c2[2] = {0, 0};
                        world_hit_legacy(sc, ro, rd, p->absorptivity, &h, count ? c2 : NULL);
                        tot_nodes += c2[0];
                        tot_prims += c2[1];
                        if (h.t >= 0.0f) { /* propagate_once :980-989 */
                            if (vdot(rd, h.normal) > 0.0f) {
                                h.normal = vneg(h.normal);
                                h.ior = 1.0f / h.ior;
                                h.absorptivity = 0.0f;
                            }
                            scatter_legacy(&h, &ro, &rd, &l, u, u2);
                        } else { /* :990-991 */
                            V3 e = sc->env ? environment_color(sc, rd) : background_color(rd);
                            radiance = vmul(e, l);
                            ended = 1;
                            break;
                        }
                    } else if (p->shading_model == PT_SHADE_LEGACY_STAGE6 || p->shading_model == PT_SHADE_LEGACY_STAGE7) {
                        float t;
                        int id = world_hit_stage(sc->sph_cr, sc->n_sph, ro, rd, &t);
                        tot_prims += (uint64_t)sc->n_sph;
                        if (id >= 0) {
                            V3 point = vadd(ro, vscale(rd, t));
                            V3 normal = vnormalized(vsub(point, vload(sc->sph_cr + 4 * id)));
                            scatter_stage(sc->sph_mat + id, p->shading_model, p->absorptivity, point, normal, &ro, &rd, &l, u, u2);
                        } else {
                            radiance = vmul(background_color(rd), l);
                            ended = 1;
                            break;
                        }
                    } else {
                        float t;
                        int id = world_hit_v2(sc->sph_cr, sc->sph_mat, sc->n_sph, ro, rd, &t);
                        tot_prims += (uint64_t)sc->n_sph;
                        if (id >= 0) { /* propagate_once __main__.py:65-75 */
                            const PtMaterial* m = sc->sph_mat + id;
                            V3 point = vadd(ro, vscale(rd, t));                       /* world.py:57 */
                            V3 normal = vnormalized(vsub(point, vload(sc->sph_cr + 4 * id))); /* :58 */
                            float ior = m->ior;
                            if (p->shading_model == PT_SHADE_V2_NORMALS) { /* 5_anti_aliasing/__main__.py:19-28 */
                                radiance = v3(0.5f * (normal.x + 1.0f), 0.5f * (normal.y + 1.0f), 0.5f * (normal.z + 1.0f));
                                ended = 1;
                                break;
                            }
                            if (vdot(rd, normal) > 0.0f) { /* world.py:31-33 */
                                normal = vneg(normal);
                                ior = 1.0f / ior;
                            }
                            if (p->shading_model == PT_SHADE_V2_DIFFUSE) { /* 6_diffuse/bsdf.py:20-26 */
                                l = vmul(l, vload(m->albedo));
                                ro = point;
                                rd = sample_lambertian(normal, u[0], u[1]);
                            } else {
                                scatter_v2(m, ior, point, normal, &ro, &rd, &l, u, u2);
                            }
                        } else {
                            radiance = vmul(background_color(rd), l); /* __main__.py:86-87 */
                            ended = 1;
                            break;
                        }
                    }
                }
                if (ended && finite3(radiance)) {
                    size_t o = ((size_t)i * H + j) * 3;
                    accum[o] += radiance.x; accum[o + 1] += radiance.y; accum[o + 2] += radiance.z;
                    if (accum_sq) {
                        accum_sq[o] += radiance.x * radiance.x;
                        accum_sq[o + 1] += radiance.y * radiance.y;
                        accum_sq[o + 2] += radiance.z * radiance.z;
                    }
                }
            }
        }
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->paths = (uint64_t)W * H * (uint64_t)p->spp;
        stats->segments = tot_seg;
        stats->nodes_visited = tot_nodes;
        stats->prims_tested = tot_prims;
        stats->iterations = p->spp;
    }
    return PT_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* post — ref: v2 postprocessing.py:5-29, __main__.py:90-96; legacy:1016-1019                  */
/* ------------------------------------------------------------------------------------------ */
void orc_postprocess(const float* accum, int width, int height, float scale, int aces, float gamma, float* out) {
    static const float mi[9] = {0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f};
    static const float mo[9] = {1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f};
    const size_t n = (size_t)width * height;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) {
        float c[3] = {accum[3 * k] * scale, accum[3 * k + 1] * scale, accum[3 * k + 2] * scale};
        if (aces) {
            float v[3], w[3];
            for (int r = 0; r < 3; ++r) v[r] = (mi[3 * r] * c[0] + mi[3 * r + 1] * c[1]) + mi[3 * r + 2] * c[2];
            for (int r = 0; r < 3; ++r) {
                float a = v[r] * (v[r] + 0.0245786f) - 0.000090537f;
                float b = v[r] * (0.983729f * v[r] + 0.4329510f) + 0.238081f;
                w[r] = a / b;
            }
            for (int r = 0; r < 3; ++r)
                c[r] = fmaxf((mo[3 * r] * w[0] + mo[3 * r + 1] * w[1]) + mo[3 * r + 2] * w[2], 0.0f);
        }
        for (int r = 0; r < 3; ++r) out[3 * k + r] = powf(c[r], 1.0f / gamma);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* raw-triangle batches                                                   */
/* ------------------------------------------------------------------------------------------ */
void orc_trace_triangles(const float* tris, int64_t ntri, const float* rays, int64_t nrays, int32_t* prim_id,
                         float* t, float* t_second) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < nrays; ++k) {
        V3 o = vload(rays + 8 * k), d = vload(rays + 8 * k + 4);
        float bt = -1.0f, bt2 = -1.0f, w[3];
        int id = -1;
        for (int64_t f = 0; f < ntri; ++f) {
            const float* T = tris + 9 * f;
            float tt = triangle_hit_t(vload(T), vload(T + 3), vload(T + 6), o, d, w);
            if (tt > ORC_EPS) {
                if (bt < 0.0f || tt < bt) { bt2 = bt; bt = tt; id = (int)f; }
                else if (bt2 < 0.0f || tt < bt2) bt2 = tt;
            }
        }
        prim_id[k] = id;
        t[k] = bt;
        if (t_second) t_second[k] = bt2;
    }
}

/* ordered, pruned traversal over a downloaded BVH2 (layout: pt_scene_bvh_download).  The leaf test is
 * the REFERENCE's plane+barycentric test on the original vertices (tris9), so the result equals the
 * brute-force orc_trace_triangles result whenever the boxes are conservative. */
void orc_trace_bvh2(const float* nodes, int64_t n_nodes, const float* tris9, int64_t ntri, const float* rays,
                    int64_t nrays, int32_t* prim_id, float* t, uint64_t counts[2]) {
    uint64_t cn = 0, ct = 0;
    (void)ntri;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cn, ct)
    for (int64_t k = 0; k < nrays; ++k) {
        V3 o = vload(rays + 8 * k), d = vload(rays + 8 * k + 4);
        V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float bt = INFINITY, w[3];
        int id = -1;
        int32_t stack[128];
        int sp = 0;
        if (n_nodes > 0) stack[sp++] = 0;
        while (sp > 0) {
            int32_t cur = stack[--sp];
            if (cur < 0) { /* leaf */
                int f = ~cur;
                const float* T = tris9 + 9 * (size_t)f;
                ct++;
                float tt = triangle_hit_t(vload(T), vload(T + 3), vload(T + 6), o, d, w);
                if (tt > ORC_EPS && (tt < bt || (tt == bt && f < id))) { bt = tt; id = f; }
                continue;
            }
            cn++;
            const float* N = nodes + 16 * (size_t)cur;
            float tn[2];
            int hit[2];
            for (int c = 0; c < 2; ++c) {
                const float* lo = N + 6 * c;
                const float* hi = lo + 3;
                float ax = (lo[0] - o.x) * inv.x, bx = (hi[0] - o.x) * inv.x;
                float ay = (lo[1] - o.y) * inv.y, by = (hi[1] - o.y) * inv.y;
                float az = (lo[2] - o.z) * inv.z, bz = (hi[2] - o.z) * inv.z;
                float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
                float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
                /* conservative: widen by a relative 1e-5 so rounding never culls the reference hit */
                hit[c] = t1 * 1.00001f + 1e-6f >= t0 && t0 * 0.99999f - 1e-6f <= bt;
                tn[c] = t0;
            }
            int32_t c0, c1;
            memcpy(&c0, N + 12, 4);
            memcpy(&c1, N + 13, 4);
            if (hit[0] && hit[1]) {
                if (tn[0] <= tn[1]) { stack[sp++] = c1; stack[sp++] = c0; }
                else { stack[sp++] = c0; stack[sp++] = c1; }
            } else if (hit[0]) stack[sp++] = c0;
            else if (hit[1]) stack[sp++] = c1;
        }
        prim_id[k] = id;
        t[k] = id >= 0 ? bt : -1.0f;
    }
    if (counts) { counts[0] = cn; counts[1] = ct; }
}

/* the same walk over a 4-wide tree (layout: learn_path_tracing_b200/csrc/bvh4.h — minx[4] miny[4] minz[4] maxx[4] maxy[4]
 * maxz[4] ref[4] pad[4]; ref 0x7fffffff = empty slot): nearest hit child first, the others in slot order, like the device
 * step.  Checks the collapsed tree: the hits must equal orc_trace_bvh2's on the tree it was collapsed from. */
void orc_trace_bvh4(const float* wnodes, int64_t n_nodes, const float* tris9, int64_t ntri, const float* rays,
                    int64_t nrays, int32_t* prim_id, float* t, uint64_t counts[2]) {
    uint64_t cn = 0, ct = 0;
    (void)ntri;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cn, ct)
    for (int64_t k = 0; k < nrays; ++k) {
        V3 o = vload(rays + 8 * k), d = vload(rays + 8 * k + 4);
        V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float bt = INFINITY, w[3];
        int id = -1;
        int32_t stack[256];
        int sp = 0;
        if (n_nodes > 0) stack[sp++] = 0;
        while (sp > 0) {
            int32_t cur = stack[--sp];
            if (cur < 0) { /* leaf */
                int f = ~cur;
                const float* T = tris9 + 9 * (size_t)f;
                ct++;
                float tt = triangle_hit_t(vload(T), vload(T + 3), vload(T + 6), o, d, w);
                if (tt > ORC_EPS && (tt < bt || (tt == bt && f < id))) { bt = tt; id = f; }
                continue;
            }
            cn++;
            const float* N = wnodes + 32 * (size_t)cur;
            int32_t ref[4];
            memcpy(ref, N + 24, 16);
            float tn[4];
            int hit[4], near = -1;
            for (int c = 0; c < 4; ++c) {
                float ax = (N[c] - o.x) * inv.x, bx = (N[12 + c] - o.x) * inv.x;
                float ay = (N[4 + c] - o.y) * inv.y, by = (N[16 + c] - o.y) * inv.y;
                float az = (N[8 + c] - o.z) * inv.z, bz = (N[20 + c] - o.z) * inv.z;
                float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
                float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
                hit[c] = ref[c] != 0x7fffffff && t1 * 1.00001f + 1e-6f >= t0 && t0 * 0.99999f - 1e-6f <= bt;
                tn[c] = t0;
                if (hit[c] && (near < 0 || t0 < tn[near])) near = c;
            }
            for (int c = 3; c >= 0; --c)
                if (hit[c] && c != near) stack[sp++] = ref[c];
            if (near >= 0) stack[sp++] = ref[near];
        }
        prim_id[k] = id;
        t[k] = id >= 0 ? bt : -1.0f;
    }
    if (counts) { counts[0] = cn; counts[1] = ct; }
}

/* reference triangle test for given (ray, triangle) pairs: t (or -1) and the smallest barycentric */
void orc_triangle_eval(const float* tris9, const int32_t* ids, const float* rays, int64_t nrays, float* t,
                       float* wmin) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nrays; ++k) {
        if (ids[k] < 0) { t[k] = -1.0f; wmin[k] = -1.0f; continue; }
        const float* T = tris9 + 9 * (size_t)ids[k];
        float w[3];
        V3 p1 = vload(T), p2 = vload(T + 3), p3 = vload(T + 6), o = vload(rays + 8 * k), d = vload(rays + 8 * k + 4);
        V3 N = vnormalized(vcross(vsub(p2, p1), vsub(p3, p1)));
        float tt = (vdot(N, p1) - vdot(o, N)) / vdot(d, N);
        triangle_hit_t(p1, p2, p3, o, d, w);
        t[k] = tt;
        wmin[k] = fminf(w[0], fminf(w[1], w[2]));
    }
}

void orc_random_triangles(int64_t n, uint32_t seed, float s, float* tris9) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        float a[4], b[4], c[4];
        orc_rng4((uint32_t)k, 0u, 0u, seed, a);
        orc_rng4((uint32_t)k, 0u, 1u, seed, b);
        orc_rng4((uint32_t)k, 0u, 2u, seed, c);
        float e1[3], e2[3], p0[3];
        for (int i = 0; i < 3; ++i) {
            e1[i] = s * (2.0f * b[i] - 1.0f);
            e2[i] = s * (2.0f * c[i] - 1.0f);
            p0[i] = a[i] - (e1[i] + e2[i]) * (1.0f / 3.0f);
        }
        float* T = tris9 + 9 * k;
        for (int i = 0; i < 3; ++i) { T[i] = p0[i]; T[3 + i] = p0[i] + e1[i]; T[6 + i] = p0[i] + e2[i]; }
    }
}

void orc_random_rays(int64_t n, uint32_t seed, float* rays8) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        float a[4], b[4];
        orc_rng4((uint32_t)k, 1u, 0u, seed, a);
        orc_rng4((uint32_t)k, 1u, 1u, seed, b);
        V3 s = sample_at_sphere(a[0], a[1]);
        V3 o = v3(0.5f + 1.5f * s.x, 0.5f + 1.5f * s.y, 0.5f + 1.5f * s.z);
        V3 d = vnormalized(vsub(v3(a[2], a[3], b[0]), o));
        float* r = rays8 + 8 * k;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = ORC_EPS;
        r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = INFINITY;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
