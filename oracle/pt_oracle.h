/*
 * pt_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the light-transport hot path of JeffreyXiang/learn_path_tracing
 * (taichi_pathtracer/10_final and legacy/PT_in_one_weekend/15_module.py), used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * learn_path_tracing_b200/ may import, link or execute it.
 *
 * PARITY PINNING: the reference is Python + Taichi and Taichi is not installable here, so the
 * reference itself cannot be executed.  The oracle is pinned against the reference's committed renders
 * outputs/{6_diffuse,7_reflect,8_refract,9_dof}.png (down-sampled copies under tests/golden/, see
 * tests/golden/make_goldens.py) and against the .world.npy <-> .obj known-answer relation.  Taichi's
 * ti.random streams and transcendental ulps are NOT pinned by anything in the reference
 * ("parity unpinned" for those): image parity is therefore statistical (3 sigma), hit parity exact.
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include <stdint.h>
#include "../include/pt_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one legacy mesh: faces are in the order of the reference's tree_leaves_field (leaf order);
 * the stored SAH tree (tree_nodes_field / tree_leaves_field_cut) is optional. */
typedef struct OrcMesh {
    const float* pos;     /* [nv][3] */
    const float* nrm;     /* [nn][3] */
    const float* uv;      /* [nt][2] */
    const int32_t* faces; /* [nf][10] = a.p,a.n,a.t,b.p,b.n,b.t,c.p,c.n,c.t,texture_id */
    int32_t nf;
    int32_t n_nodes;            /* 0: no stored tree -> brute force over faces */
    const int32_t* node_left;   /* [n_nodes] */
    const int32_t* node_right;  /* [n_nodes] */
    const float* node_low;      /* [n_nodes][3] */
    const float* node_high;     /* [n_nodes][3] */
    const int32_t* node_data;   /* [n_nodes] leaf index or -1 */
    const int32_t* leaf_cut;    /* [n_leaves+1] CSR offsets into faces */
    int32_t max_depth;
    int32_t _pad;
} OrcMesh;

typedef struct OrcScene {
    /* v2 spheres */
    const float* sph_cr;         /* [n_sph][4] */
    const PtMaterial* sph_mat;   /* [n_sph]    */
    int32_t n_sph;
    /* legacy textured spheres */
    int32_t n_tsph;
    const float* tsph_cr;        /* [n_tsph][4] */
    const int32_t* tsph_transparency;
    const int32_t* tsph_tex;
    /* legacy meshes */
    const OrcMesh* meshes;
    int32_t n_mesh;
    /* texture atlas, 8 bytes/texel, x-major [W][H] (same bytes as pt_scene_set_texture_atlas) */
    int32_t tex_W;
    const uint8_t* texels;
    const int32_t* tex_areas;    /* [ntex][4] */
    const int32_t* tex_flags;    /* [ntex] bit 0: no normal map -> (0,0,1) */
    int32_t tex_H, ntex;
    /* environment, float rgb x-major [W][H][3]; NULL -> v2 sky gradient */
    const float* env;
    int32_t env_W, env_H;
    int32_t env_area[4];
} OrcScene;

/* counter-based RNG shared (as arithmetic, not as code) with the CUDA library */
void orc_rng4(uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t seed, float out[4]);

/* Camera.get_rays (10_final/camera.py:71-93): rays[h*w][8] = o.xyz,tmin,d.xyz,tmax */
void orc_generate_rays(const PtCamera* cam, int width, int height, int sample, uint32_t seed, float* rays);

/* World.hit (10_final/world.py:24-34) over a ray batch, brute force, reference f32 arithmetic.
 * t64 (optional) = same roots evaluated in double for conditioning reports. */
void orc_trace_spheres(const float* cr, const PtMaterial* mats, int n, const float* rays, int64_t nrays,
                       int32_t* prim_id, float* t, double* t64);

/* brute-force closest hit over raw triangles tris[n][9]=p0,p1,p2 with the reference's
 * plane + barycentric test (15_module.py:909-928); t2/id2 optional = second closest distinct hit. */
void orc_trace_triangles(const float* tris, int64_t ntri, const float* rays, int64_t nrays, int32_t* prim_id,
                         float* t, float* t_second);

/* ordered, t-pruned traversal of a BVH2 in the pt_scene_bvh_download layout with the Moller-Trumbore
 * test over tris12[n][12]; counts (optional) = {nodes visited, triangles tested} totals. */
void orc_trace_bvh4(const float* wnodes, int64_t n_nodes, const float* tris9, int64_t ntri, const float* rays,
                    int64_t nrays, int32_t* prim_id, float* t, uint64_t counts[2]);
void orc_trace_bvh2(const float* nodes, int64_t n_nodes, const float* tris12, int64_t ntri, const float* rays,
                    int64_t nrays, int32_t* prim_id, float* t, uint64_t counts[2]);

/* the reference triangle test for given (ray, triangle id) pairs: plane t and smallest barycentric */
void orc_triangle_eval(const float* tris9, const int32_t* ids, const float* rays, int64_t nrays, float* t,
                       float* wmin);

/* full legacy World.hit (15_module.py:838-848) over a ray batch: global prim ids are
 * textured spheres first, then mesh faces in mesh order. */
void orc_trace_legacy(const OrcScene* sc, const float* rays, int64_t nrays, int32_t* prim_id, float* t);

/* render(): accum[w][h][3] += radiance (Taichi field layout); accum_sq optional; stats optional.
 * threads <= 0: all OpenMP threads. */
int orc_render(const OrcScene* sc, const PtCamera* cam, const PtRenderParams* p, float* accum, float* accum_sq,
               PtStats* stats, int threads);

/* post_processing (10_final/postprocessing.py:5-29) / legacy gamma_correction (15_module.py:1016-1019) */
void orc_postprocess(const float* accum, int width, int height, float scale, int aces, float gamma, float* out);

/* deterministic generators shared (as arithmetic) with the CUDA library: SURVEY 8d config 5 */
void orc_random_triangles(int64_t n, uint32_t seed, float edge_scale, float* tris9);
void orc_random_rays(int64_t n, uint32_t seed, float* rays8);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
