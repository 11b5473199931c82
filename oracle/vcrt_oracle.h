/* vcrt_oracle.h -- interface of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY (see vcrt_oracle.c).
 * Shares only the data ABI and the parameter block with the product (include/vcrt.h). */
#ifndef VCRT_ORACLE_H
#define VCRT_ORACLE_H
#include "../include/vcrt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const vcrt_triangle* triangles; uint32_t num_triangles;
    const vcrt_material* materials; uint32_t num_materials;
    const vcrt_bvh_node* bvh;       uint32_t num_bvh_nodes;
    const vcrt_light*    lights;    uint32_t num_lights;
    const vcrt_sphere*   spheres;   uint32_t num_spheres;
} vcrt_oracle_scene;

typedef struct {
    uint64_t rays;             /* closest-hit queries */
    uint64_t ref_nodes;        /* node records fetched by the literal reference traversal */
    uint64_t ref_triangles;
    uint64_t canon_nodes;      /* node records fetched by the canonical (reference order + t-culling) traversal; */
    uint64_t canon_triangles;  /*   only with VCRT_FLAG_COUNT_TRAVERSAL: defines the roofline's bytes per ray */
    uint32_t max_stack;        /* deepest stackIndex the literal traversal reached */
    uint32_t _pad;
} vcrt_oracle_counters;

/* Renders samples [sample_begin, sample_begin+sample_count) of every covered pixel.
 * VCRT_ACCUM_RGBA8_REF: target/accum8 (W*H*4 bytes each) carry the running mean exactly as the reference's
 * dispatch + image copy does.  VCRT_ACCUM_F32: accumf (W*H*4 floats) += (r,g,b,1) per sample.
 * aov may be NULL.  counters are added to (not reset). */
int vcrt_oracle_render(const vcrt_oracle_scene* scene, const vcrt_ubo* ubo, const vcrt_render_params* params,
                       uint32_t width, uint32_t height, uint8_t* target, uint8_t* accum8, float* accumf,
                       vcrt_aov* aov, vcrt_oracle_counters* counters);

int vcrt_oracle_hit_bvh(const vcrt_oracle_scene* scene, uint32_t stack_depth, const float* org_dir6, int n,
                        uint32_t* out10, int32_t* tri_out);
/* post-process-shader.frag:26-70: mix * smartDeNoise(sigma, kSigma, threshold) + (1-mix) * texel, then pow(rgb, 1/gamma)
 * (gamma <= 0: none), alpha 1.  tex/out: rgba8 W*H. */
int vcrt_oracle_post_process(const uint8_t* tex, uint32_t width, uint32_t height, float mix, float sigma, float kSigma, float threshold,
                             float gamma, uint8_t* out);
void vcrt_oracle_random(uint32_t seed, int n, float* out);
uint32_t vcrt_oracle_pcg_next(uint32_t* state);
void vcrt_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void vcrt_oracle_sincos_portable(float x, float* s, float* c);

#ifdef __cplusplus
}
#endif
#endif
