/*
 * vcrt_oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA kernels.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product (libvcrt.so and the
 * Python/C++ host layers) never links, imports or falls back to it.
 *
 * PINNED: not by reference tests (the reference has none) but by the reference itself
 * run here -- oracle/_ref/libvcrt_ref.so is the reference's own shader text compiled as C++
 * (oracle/ref/Makefile); tests/test_oracle_vs_ref.py requires this file to reproduce its rgba8 frames,
 * hit records and RNG stream bit-for-bit, and tests/golden/ holds vectors generated from it.
 *
 * What is restated (file:line in /root/reference/resources/shaders/source unless noted):
 *   main                  ray-trace-compute.comp:352-380      ray generation + running mean
 *   ray_color             ray-trace-compute.comp:314-350      path loop
 *   hit_bvh               ray-trace-compute.comp:263-311      stack traversal (+ a t-culled canonical twin for counting)
 *   intersectAABB         ray-trace-compute.comp:250-258
 *   hit_triangle          ray-trace-compute.comp:205-220, triIntersect :157-173
 *   hit_sphere/hit_scene  ray-trace-compute.comp:175-203, :222-247   (dead on the shipped path; VCRT_TRAVERSAL_BRUTE_FORCE)
 *   scatter               ray-trace-compute.comp:118-155; simple variant ray-trace-compute-simple.comp:62-68
 *   sampleLight & co      ray-trace-compute.comp:67-116
 *   Onb / onbLocal        include/definitions.glsl:42-53
 *   PCG RNG, hemisphere   include/random.glsl:4-52
 * Canonical floating point (GLSL leaves it open; these are the choices oracle/_ref makes through glm 0.9.9.9):
 *   no FMA contraction; IEEE fp32 / and sqrt; dot = (x+y)+z; normalize = v * (1/sqrt(dot(v,v)));
 *   min(x,y) = (y<x)?y:x, max(x,y) = (x<y)?y:x; reflect = I - N*dot(N,I)*2; libm sinf/cosf/tanf
 *   (or the portable sequence below when trig_mode = VCRT_TRIG_PORTABLE); unorm8 store rounds half-even.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared (oracle/Makefile).
 */
#include "vcrt_oracle.h"

#include <math.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline v3 add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 divv(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* (x+y)+z */
static inline v3 cross(v3 x, v3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
static inline v3 normalize(v3 v) { return scale(v, 1.0f / sqrtf(dot(v, v))); }
static inline float minf_(float x, float y) { return (y < x) ? y : x; }
static inline float maxf_(float x, float y) { return (x < y) ? y : x; }
static inline v3 reflect(v3 I, v3 N) { return sub(I, scale(scale(N, dot(N, I)), 2.0f)); }
static inline v3 refract(v3 I, v3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k >= 0.0f) return sub(scale(I, eta), scale(N, eta * d + sqrtf(k)));
    return V3(0.0f, 0.0f, 0.0f);
}

/* ---------------------------------------------------------------- portable sin/cos (VCRT_TRIG_PORTABLE)
 * A fixed sequence of IEEE fp32 +,-,* (no FMA): bit-identical under gcc -ffp-contract=off and nvcc -fmad=false.
 * Valid for the only argument range the path produces, phi = 2*pi*r1 in [0, 2*pi]. Cephes-style
 * three-constant Cody-Waite reduction to [-pi/4, pi/4] and degree-7/8 minimax polynomials. */
void vcrt_oracle_sincos_portable(float x, float* s, float* c) {
    int q = (int)(x * 0.636619772367581343f + 0.5f); /* nearest multiple of pi/2 (x >= 0) */
    float fq = (float)q;
    float r = ((x - fq * 1.5703125f) - fq * 4.837512969970703125e-4f) - fq * 7.549789948768648e-8f;
    float z = r * r;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (q & 3) {
        case 0: *s = ps;  *c = pc;  break;
        case 1: *s = pc;  *c = -ps; break;
        case 2: *s = -ps; *c = -pc; break;
        default: *s = -pc; *c = ps; break;
    }
}

/* ---------------------------------------------------------------- RNG */
typedef struct {
    uint32_t mode;
    uint32_t pcg;                 /* random.glsl:19 state */
    uint32_t key[2], ctr[3], buf[4], have; /* Philox: key (pixel, seed), counter (sample, bounce, block-in-bounce) */
} rng_t;

uint32_t vcrt_oracle_pcg_next(uint32_t* state) { /* random.glsl:4-17, returns `word` */
    *state = *state * 747796405u + 1u;
    uint32_t s = *state;
    uint32_t word = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (word >> 22) ^ word;
}

void vcrt_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float rng_next(rng_t* g) {
    if (g->mode == VCRT_RNG_PCG_REF) {
        uint32_t w = vcrt_oracle_pcg_next(&g->pcg);
        return (float)w / 4294967295.0f; /* float(2^32-1) == 2^32: [0,1] inclusive, random.glsl:16 */
    }
    if (g->have == 0) {
        uint32_t c[4] = {g->ctr[0], g->ctr[1], g->ctr[2], 0u};
        vcrt_oracle_philox4x32_10(c, g->key, g->buf);
        g->ctr[2]++;
        g->have = 4;
    }
    uint32_t w = g->buf[4 - g->have];
    g->have--;
    return (float)(w >> 8) * 5.9604644775390625e-8f; /* [0,1) */
}

/* ---------------------------------------------------------------- scene access with robustBufferAccess semantics */
typedef struct {
    const vcrt_oracle_scene* s;
    uint32_t lights_length, stack_depth, max_bounces, shader, traversal, trig;
    uint32_t ubo_num_triangles;   /* ubo.numTriangles: loop bound of hit_scene (:229) */
    int count;
} env_t;

static const vcrt_triangle ZERO_TRI;
static const vcrt_material ZERO_MAT;
static const vcrt_light ZERO_LIGHT;
static const vcrt_bvh_node ZERO_NODE;

static inline const vcrt_triangle* tri_at(const env_t* e, int64_t i) { return (i < 0 || i >= (int64_t)e->s->num_triangles) ? &ZERO_TRI : &e->s->triangles[i]; }
static inline const vcrt_material* mat_at(const env_t* e, int64_t i) { return (i < 0 || i >= (int64_t)e->s->num_materials) ? &ZERO_MAT : &e->s->materials[i]; }
static inline const vcrt_light* light_at(const env_t* e, int64_t i) { return (i < 0 || i >= (int64_t)e->s->num_lights) ? &ZERO_LIGHT : &e->s->lights[i]; }
static inline const vcrt_bvh_node* node_at(const env_t* e, int64_t i) { return (i < 0 || i >= (int64_t)e->s->num_bvh_nodes) ? &ZERO_NODE : &e->s->bvh[i]; }

typedef struct { v3 origin, dir; } ray_t;
typedef struct { v3 p, normal; uint32_t materialIndex; float t; int backFaceInt; int triangle; } hit_t;

typedef struct { uint64_t rays, ref_nodes, ref_tris, canon_nodes, canon_tris; uint32_t max_stack; } tally_t;

/* ray-trace-compute.comp:205-220 + :157-173 */
static inline int hit_triangle(const env_t* e, int ti, const ray_t* r, float tMin, float tMax, hit_t* rec) {
    const vcrt_triangle* tr = tri_at(e, ti);
    v3 v0 = ld3(tr->v0), v1 = ld3(tr->v1), v2 = ld3(tr->v2);
    v3 a = sub(v0, v1), b = sub(v2, v0), p = sub(v0, r->origin);
    v3 n = cross(b, a);
    v3 q = cross(p, r->dir);
    float idet = 1.0f / dot(r->dir, n);
    float u = dot(q, b) * idet, v = dot(q, a) * idet, t = dot(n, p) * idet;
    if (!(u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f)) {
        rec->p = add(r->origin, scale(r->dir, t));
        rec->normal = normalize(n);
        rec->backFaceInt = dot(r->dir, rec->normal) > 0.0f ? 1 : 0;
        rec->normal = scale(rec->normal, (float)(1 - 2 * rec->backFaceInt));
        rec->p = add(rec->p, scale(rec->normal, 0.0001f));
        rec->t = t;
        rec->materialIndex = tr->materialIndex;
        rec->triangle = ti;
        return t > tMin && t < tMax;
    }
    return 0;
}

/* ray-trace-compute.comp:175-203 */
static inline int hit_sphere(const env_t* e, int si, const ray_t* r, float tMin, float tMax, hit_t* rec) {
    const vcrt_sphere* sp = &e->s->spheres[si];
    v3 center = V3(sp->s[0], sp->s[1], sp->s[2]);
    float radius = sp->s[3];
    v3 oc = sub(r->origin, center);
    float a = dot(r->dir, r->dir), half_b = dot(oc, r->dir), c = dot(oc, oc) - radius * radius;
    float disc = half_b * half_b - a * c;
    if (disc < 0.0f) return 0;
    float sqrtd = sqrtf(disc);
    rec->backFaceInt = 0;
    float root = (-half_b - sqrtd) / a;
    if (root < tMin || tMax < root) {
        root = (-half_b + sqrtd) / a;
        rec->backFaceInt = 1;
        if (root < tMin || tMax < root) return 0;
    }
    rec->t = root;
    rec->p = add(r->origin, scale(r->dir, root));
    /* (1 - 2*bf) * (p - center) / radius : int*vec3 first, then / radius */
    v3 d = scale(sub(rec->p, center), (float)(1 - 2 * rec->backFaceInt));
    rec->normal = V3(d.x / radius, d.y / radius, d.z / radius);
    rec->materialIndex = sp->materialIndex;
    rec->triangle = -2 - si;
    return 1;
}

/* ray-trace-compute.comp:222-247 (brute force; the sphere loop restarts from t_max, so a sphere hit overrides a nearer triangle) */
static int hit_scene(const env_t* e, const ray_t* r, hit_t* rec) {
    const float t_min = 0.001f, t_max = 10000.0f;
    hit_t tmp;
    int hit_anything = 0;
    float closest = t_max;
    for (uint32_t i = 0; i < e->ubo_num_triangles; ++i)   /* :229 -- ubo.numTriangles, reads past the buffer return the zero triangle */
        if (hit_triangle(e, (int)i, r, t_min, closest, &tmp)) { hit_anything = 1; closest = tmp.t; *rec = tmp; }
    if (e->shader == VCRT_SHADER_SIMPLE) return hit_anything;   /* ray-trace-compute-simple.comp:106-123 has no sphere loop */
    closest = t_max;
    for (uint32_t j = 0; j < e->s->num_spheres; ++j)
        if (hit_sphere(e, (int)j, r, t_min, closest, &tmp)) { hit_anything = 1; closest = tmp.t; *rec = tmp; }
    return hit_anything;
}

/* ray-trace-compute.comp:250-258 */
static inline void intersect_aabb(const ray_t* r, const vcrt_bvh_node* nd, float* tNear, float* tFar) {
    v3 tMin = divv(sub(ld3(nd->min), r->origin), r->dir);
    v3 tMax = divv(sub(ld3(nd->max), r->origin), r->dir);
    v3 t1 = V3(minf_(tMin.x, tMax.x), minf_(tMin.y, tMax.y), minf_(tMin.z, tMax.z));
    v3 t2 = V3(maxf_(tMin.x, tMax.x), maxf_(tMin.y, tMax.y), maxf_(tMin.z, tMax.z));
    *tNear = maxf_(maxf_(t1.x, t1.y), t1.z);
    *tFar = minf_(minf_(t2.x, t2.y), t2.z);
}

#define ORACLE_MAX_STACK 64

/* ray-trace-compute.comp:263-311.  culled = 0: literal; culled = 1: the canonical t-culled twin whose
 * fetch counts define the roofline's bytes per ray; it returns the same hit. */
static int hit_bvh(const env_t* e, const ray_t* r, hit_t* rec, int culled, uint64_t* nodes, uint64_t* tris, uint32_t* max_sp) {
    const float t_min = 0.001f, t_max = 10000.0f;
    const int depth = (int)e->stack_depth;
    int hit_anything = 0;
    float closest = t_max;
    int stack[ORACLE_MAX_STACK];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0 && sp < depth) {
        sp--;
        int cur = stack[sp];
        if (cur == -1) continue;
        const vcrt_bvh_node* nd = node_at(e, cur);
        ++*nodes;
        float tN, tF;
        intersect_aabb(r, nd, &tN, &tF);
        if (tN > tF) continue;
        if (culled && (tF < t_min || tN > closest)) continue;
        int ti = nd->objectIndex;
        if (ti != -1) {
            hit_t tmp;
            ++*tris;
            if (hit_triangle(e, ti, r, t_min, closest, &tmp)) { hit_anything = 1; closest = tmp.t; *rec = tmp; }
        }
        stack[sp++] = nd->leftNodeIndex;
        stack[sp++] = nd->rightNodeIndex;
        if ((uint32_t)sp > *max_sp) *max_sp = (uint32_t)sp;
    }
    return hit_anything;
}

static int closest_hit(const env_t* e, const ray_t* r, hit_t* rec, tally_t* tl) {
    tl->rays++;
    if (e->traversal == VCRT_TRAVERSAL_BRUTE_FORCE) return hit_scene(e, r, rec);
    int h = hit_bvh(e, r, rec, 0, &tl->ref_nodes, &tl->ref_tris, &tl->max_stack);
    if (e->count) {
        hit_t rec2; uint32_t ms = 0;
        memset(&rec2, 0, sizeof rec2);
        int h2 = hit_bvh(e, r, &rec2, 1, &tl->canon_nodes, &tl->canon_tris, &ms);
        (void)h2;
    }
    return h;
}

/* include/definitions.glsl:42-53 + include/random.glsl:42-52 + ray-trace-compute.comp:91-99 */
static v3 sample_lambertian(const env_t* e, v3 normal, rng_t* g) {
    v3 w = normalize(normal);
    v3 a = (fabsf(w.x) > 0.9f) ? V3(0, 1, 0) : V3(1, 0, 0);
    v3 v = normalize(cross(w, a));
    v3 u = cross(w, v);
    float r1 = rng_next(g), r2 = rng_next(g);
    float z = sqrtf(1.0f - r2);
    float phi = 2.0f * 3.1415926535897932385f * r1;
    float sn, cs;
    if (e->trig == VCRT_TRIG_PORTABLE) vcrt_oracle_sincos_portable(phi, &sn, &cs);
    else { cs = cosf(phi); sn = sinf(phi); }
    float sr2 = sqrtf(r2);
    float x = cs * sr2, y = sn * sr2;
    return normalize(add(add(scale(u, x), scale(v, y)), scale(w, z)));
}

/* ray-trace-compute.comp:106-116 */
static v3 sample_glass(v3 I, const hit_t* rec) {
    float ir = 1.5f;
    float ratio = (float)(1 - rec->backFaceInt) * 1.0f / ir + (float)rec->backFaceInt * ir;
    v3 i = normalize(I);
    float cos_theta = minf_(dot(neg(i), rec->normal), 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float t = floorf(minf_(maxf_(ratio * sin_theta, 0.0f), 1.0f));
    return add(scale(reflect(i, rec->normal), t), scale(refract(i, rec->normal, ratio), 1.0f - t));
}

/* ray-trace-compute.comp:67-89 */
static v3 sample_light(const env_t* e, v3 p, rng_t* g, float* lightCosine) {
    int lightIndex = (int)floorf((float)(int)e->lights_length * rng_next(g));
    uint32_t ti = light_at(e, lightIndex)->triangleIndex;
    float s = rng_next(g), t = rng_next(g);
    const vcrt_triangle* tr = tri_at(e, ti);
    v3 v0 = ld3(tr->v0);
    v3 v01 = add(neg(v0), ld3(tr->v1)), v02 = add(neg(v0), ld3(tr->v2));
    v3 onLight = add(add(v0, scale(v01, s)), scale(v02, t));
    v3 toLight = normalize(sub(onLight, p));
    *lightCosine = fabsf(toLight.y);
    return toLight;
}

/* ray-trace-compute.comp:118-155 ; simple: ray-trace-compute-simple.comp:62-68 + random.glsl:24-40 */
static int scatter(const env_t* e, const ray_t* r_in, const hit_t* rec, v3* albedo, ray_t* scattered, rng_t* g) {
    const vcrt_material* m = mat_at(e, rec->materialIndex);
    *albedo = ld3(m->albedo);
    if (e->shader == VCRT_SHADER_SIMPLE) {
        float lo = -0.3f, hi = 0.3f;
        float px = lo + (hi - lo) * rng_next(g);
        float py = lo + (hi - lo) * rng_next(g);
        float pz = lo + (hi - lo) * rng_next(g);
        v3 p = normalize(V3(px, py, pz));
        if (!(dot(p, rec->normal) > 0.0f)) p = neg(p);
        scattered->origin = rec->p;
        scattered->dir = normalize(p);
        return m->type == VCRT_MAT_LIGHT;
    }
    v3 materialSample = V3(0, 0, 0); /* undefined in the shader for LIGHT hits; the path ends there */
    if (m->type == VCRT_MAT_LAMBERTIAN) materialSample = sample_lambertian(e, rec->normal, g);
    else if (m->type == VCRT_MAT_METAL) materialSample = reflect(r_in->dir, rec->normal);
    else if (m->type == VCRT_MAT_GLASS) { materialSample = sample_glass(r_in->dir, rec); *albedo = V3(1.0f, 1.0f, 1.0f); }
    v3 finalSample = materialSample;
    float coin = rng_next(g); /* always drawn: left operand of && (:138) */
    if (coin < 0.5f && m->type == VCRT_MAT_LAMBERTIAN) {
        float lightCosine;
        finalSample = sample_light(e, rec->p, g, &lightCosine);
        if (fabsf(lightCosine) < 0.001f) finalSample = materialSample;
    }
    scattered->origin = rec->p;
    scattered->dir = finalSample;
    return m->type == VCRT_MAT_LIGHT;
}

/* ray-trace-compute.comp:314-350 */
static v3 ray_color(const env_t* e, ray_t r, rng_t* g, tally_t* tl, vcrt_aov* aov) {
    hit_t rec;
    memset(&rec, 0, sizeof rec);
    v3 final_color = V3(1.0f, 1.0f, 1.0f);
    ray_t cur = {r.origin, normalize(r.dir)};
    for (uint32_t i = 0; i < e->max_bounces; ++i) {
        g->ctr[1] = i; g->ctr[2] = 0; g->have = 0; /* Philox: a fresh counter block per bounce (stateless across rays) */
        int hit = closest_hit(e, &cur, &rec, tl);
        if (i == 0 && aov) {
            if (hit) { aov->triangle = rec.triangle; aov->material = (int32_t)rec.materialIndex; aov->t = rec.t; aov->backFace = (uint32_t)rec.backFaceInt; }
            else { aov->triangle = -1; aov->material = -1; aov->t = 0.0f; aov->backFace = 0; }
        }
        if (hit) {
            v3 albedo;
            ray_t next;
            int emits = scatter(e, &cur, &rec, &albedo, &next, g);
            cur = next;
            final_color = mul(final_color, albedo);
            if (emits) break;
        } else {
            final_color = scale(final_color, 0.0f);
            break;
        }
    }
    return final_color;
}

static inline uint8_t unorm8(float f) {
    if (!(f == f)) return 0;
    f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
    return (uint8_t)rintf(f * 255.0f);
}

int vcrt_oracle_render(const vcrt_oracle_scene* scene, const vcrt_ubo* ubo, const vcrt_render_params* prm,
                       uint32_t W, uint32_t H, uint8_t* target, uint8_t* accum8, float* accumf, vcrt_aov* aov,
                       vcrt_oracle_counters* counters) {
    if (!scene || !ubo || !prm || W == 0 || H == 0) return VCRT_ERR_INVALID;
    env_t e;
    e.s = scene;
    e.shader = prm->shader;
    e.traversal = prm->traversal;
    e.trig = prm->trig_mode;
    e.max_bounces = prm->max_bounces ? prm->max_bounces : (prm->shader == VCRT_SHADER_SIMPLE ? 4u : 2u);
    e.stack_depth = prm->stack_depth ? prm->stack_depth : 16u;
    if (e.stack_depth > ORACLE_MAX_STACK) return VCRT_ERR_INVALID;
    e.lights_length = prm->lights_length ? prm->lights_length : scene->num_lights;
    e.count = (prm->flags & VCRT_FLAG_COUNT_TRAVERSAL) != 0;
    e.ubo_num_triangles = ubo->numTriangles;
    const uint32_t spp = prm->sample_count ? prm->sample_count : 1u;
    const uint32_t tilesX = (W + 31) / 32;
    const uint32_t covW = (prm->flags & VCRT_FLAG_REF_DISPATCH_COVERAGE) ? (W / 32) * 32 : W;
    const uint32_t covH = (prm->flags & VCRT_FLAG_REF_DISPATCH_COVERAGE) ? (H / 32) * 32 : H;
    if (prm->accum_mode == VCRT_ACCUM_RGBA8_REF && (!target || !accum8)) return VCRT_ERR_INVALID;
    if (prm->accum_mode == VCRT_ACCUM_F32 && !accumf) return VCRT_ERR_INVALID;

    /* camera, ray-trace-compute.comp:355-369 */
    const float pi = 3.1415926535897932385f;
    const float imW = (float)W, imH = (float)H;
    float vfov = 30.0f;
    float theta = vfov * pi / 180.0f;
    float hh = tanf(theta / 2.0f);
    float viewport_height = 2.0f * hh;
    float viewport_width = imW / imH * viewport_height;
    float focal_length = 1.0f;
    v3 horizontal = V3(viewport_width, 0.0f, 0.0f);
    v3 vertical = V3(0.0f, -viewport_height, 0.0f);
    v3 origin = mul(V3(ubo->camPos[2], ubo->camPos[0], ubo->camPos[1]), V3(-1.0f, 1.0f, 1.0f));
    v3 half_h = V3(horizontal.x / 2.0f, horizontal.y / 2.0f, horizontal.z / 2.0f);
    v3 half_v = V3(vertical.x / 2.0f, vertical.y / 2.0f, vertical.z / 2.0f);
    v3 llc = sub(sub(sub(origin, half_h), half_v), V3(0.0f, 0.0f, focal_length));

    tally_t total;
    memset(&total, 0, sizeof total);

#pragma omp parallel
    {
        tally_t tl;
        memset(&tl, 0, sizeof tl);
#pragma omp for schedule(dynamic, 4) nowait
        for (uint32_t y = 0; y < covH; ++y) {
            for (uint32_t x = 0; x < covW; ++x) {
                if (prm->tile_count > 1) {
                    uint32_t tile = (y / 32) * tilesX + (x / 32);
                    if (tile % prm->tile_count != prm->tile_rank) continue;
                }
                const size_t pix = (size_t)y * W + x;
                float u = (float)x / imW, v = (float)y / imH;
                ray_t r;
                r.origin = origin;
                r.dir = sub(add(add(llc, scale(horizontal, u)), scale(vertical, v)), origin);
                for (uint32_t k = 0; k < spp; ++k) {
                    const uint32_t s = prm->sample_begin + k;
                    rng_t g;
                    memset(&g, 0, sizeof g);
                    g.mode = prm->rng_mode;
                    g.pcg = (600u * x + y) * (s + 1u); /* random.glsl:19 */
                    g.key[0] = (uint32_t)pix; g.key[1] = prm->philox_seed; g.ctr[0] = s; g.ctr[1] = 0;
                    v3 c = ray_color(&e, r, &g, &tl, (aov && k == 0 && (prm->flags & VCRT_FLAG_WRITE_AOV)) ? &aov[pix] : 0);
                    if (prm->accum_mode == VCRT_ACCUM_F32) {
                        float* a = accumf + 4 * pix;
                        a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += 1.0f;
                    } else {
                        /* ray-trace-compute.comp:375-379, then the host's target -> accumulation copy (main.cpp:253-261) */
                        uint8_t* t8 = target + 4 * pix;
                        uint8_t* a8 = accum8 + 4 * pix;
                        float fs = (float)s;
                        float m = minf_(fs, 1.0f);
                        float col[4] = {c.x, c.y, c.z, 1.0f};
                        for (int ch = 0; ch < 4; ++ch) {
                            float curc = ((float)a8[ch] / 255.0f) * m;
                            float w = (col[ch] + curc * fs) / (fs + 1.0f);
                            t8[ch] = unorm8(w);
                        }
                        memcpy(a8, t8, 4);
                    }
                }
            }
        }
#pragma omp critical
        {
            total.rays += tl.rays; total.ref_nodes += tl.ref_nodes; total.ref_tris += tl.ref_tris;
            total.canon_nodes += tl.canon_nodes; total.canon_tris += tl.canon_tris;
            if (tl.max_stack > total.max_stack) total.max_stack = tl.max_stack;
        }
    }
    if (counters) {
        counters->rays += total.rays; counters->ref_nodes += total.ref_nodes; counters->ref_triangles += total.ref_tris;
        counters->canon_nodes += total.canon_nodes; counters->canon_triangles += total.canon_tris;
        if (total.max_stack > counters->max_stack) counters->max_stack = total.max_stack;
    }
    return VCRT_OK;
}

/* Closest-hit query for caller-supplied rays (dir used as given), literal reference traversal.
 * out10[i] = {hit, materialIndex, backFaceInt, t, p.xyz, normal.xyz} as raw 32-bit words; tri_out[i] = triangle or -1. */
int vcrt_oracle_hit_bvh(const vcrt_oracle_scene* scene, uint32_t stack_depth, const float* org_dir6, int n, uint32_t* out10, int32_t* tri_out) {
    env_t e;
    memset(&e, 0, sizeof e);
    e.s = scene;
    e.stack_depth = stack_depth ? stack_depth : 16u;
    if (e.stack_depth > ORACLE_MAX_STACK) return VCRT_ERR_INVALID;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        ray_t r = {ld3(org_dir6 + 6 * i), ld3(org_dir6 + 6 * i + 3)};
        hit_t rec;
        memset(&rec, 0, sizeof rec);
        uint64_t a = 0, b = 0; uint32_t ms = 0;
        int h = hit_bvh(&e, &r, &rec, 0, &a, &b, &ms);
        uint32_t* o = out10 + 10 * (size_t)i;
        memset(o, 0, 40);
        o[0] = h ? 1u : 0u;
        if (h) {
            o[1] = rec.materialIndex; o[2] = (uint32_t)rec.backFaceInt;
            memcpy(o + 3, &rec.t, 4); memcpy(o + 4, &rec.p, 12); memcpy(o + 7, &rec.normal, 12);
        }
        if (tri_out) tri_out[i] = h ? rec.triangle : -1;
    }
    return VCRT_OK;
}

void vcrt_oracle_random(uint32_t seed, int n, float* out) {
    rng_t g;
    memset(&g, 0, sizeof g);
    g.mode = VCRT_RNG_PCG_REF;
    g.pcg = seed;
    for (int i = 0; i < n; ++i) out[i] = rng_next(&g);
}

/* ---------------------------------------------------------------- post-process pass
 * post-process-shader.frag:26-70 restated: smartDeNoise (:26-60; commented out of main at :64) blended with the plain
 * texel by `mix`, then pow(rgb, 1/gamma), alpha 1 (:67-68).  PINNED: oracle/_ref compiles the fragment shader's own text
 * (as shipped, and with its commented-out denoiser line enabled) and tests/test_oracle.py requires this restatement to
 * reproduce its frames bit for bit.
 * `tex` is the rgba8 target sampled as the reference's sampler does (Image.cpp:353-364): normalised coordinates, LINEAR
 * filter, REPEAT addressing, evaluated as the Vulkan specification writes it -- texel coordinate u*W - 0.5, floor +
 * fraction, weights at 8 bits of sub-texel precision, the four taps combined in the order of the spec's formula
 * (oracle/ref/glsl_prelude.hpp: texture()).  The fragment of pixel (px, py) of the full-screen quad (mesh.cpp:58-93) has
 * fragTexCoord = ((px + 0.5) / W, (py + 0.5) / H). */
static inline void texel(const uint8_t* tex, int w, int h, int x, int y, float out[4]) {
    x %= w; if (x < 0) x += w;
    y %= h; if (y < 0) y += h;
    const uint8_t* p = tex + 4 * ((size_t)y * w + x);
    for (int c = 0; c < 4; ++c) out[c] = (float)p[c] / 255.0f;
}

static inline void texture_linear(const uint8_t* tex, int w, int h, float uvx, float uvy, float out[4]) {
    const float u = uvx * (float)w - 0.5f, v = uvy * (float)h - 0.5f;
    const float fu = floorf(u), fv = floorf(v);
    const float a = rintf((u - fu) * 256.0f) / 256.0f, b = rintf((v - fv) * 256.0f) / 256.0f;
    const int i0 = (int)fu, j0 = (int)fv;
    float t00[4], t10[4], t01[4], t11[4];
    texel(tex, w, h, i0, j0, t00); texel(tex, w, h, i0 + 1, j0, t10); texel(tex, w, h, i0, j0 + 1, t01); texel(tex, w, h, i0 + 1, j0 + 1, t11);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    for (int c = 0; c < 4; ++c) out[c] = ((w00 * t00[c] + w10 * t10[c]) + w01 * t01[c]) + w11 * t11[c];
}

int vcrt_oracle_post_process(const uint8_t* tex, uint32_t width, uint32_t height, float mix, float sigma, float kSigma, float threshold,
                             float gamma, uint8_t* out) {
    if (!tex || !out || width == 0 || height == 0) return VCRT_ERR_INVALID;
    const int w = (int)width, h = (int)height;
    const float INV_SQRT_OF_2PI = 0.39894228040143267793994605993439f, INV_PI = 0.31830988618379067153776752674503f;
#pragma omp parallel for schedule(static)
    for (int py = 0; py < h; ++py)
        for (int px = 0; px < w; ++px) {
            const float uvx = ((float)px + 0.5f) / (float)w, uvy = ((float)py + 0.5f) / (float)h;
            float centr[4];
            texture_linear(tex, w, h, uvx, uvy, centr);
            float col[4] = {centr[0], centr[1], centr[2], centr[3]};
            if (mix != 0.0f) {
                float radius = roundf(kSigma * sigma);
                float radQ = radius * radius;
                float invSigmaQx2 = 0.5f / (sigma * sigma);
                float invSigmaQx2PI = INV_PI * invSigmaQx2;
                float invThresholdSqx2 = 0.5f / (threshold * threshold);
                float invThresholdSqrt2PI = INV_SQRT_OF_2PI / threshold;
                float zBuff = 0.0f, aBuff[4] = {0, 0, 0, 0};
                for (float x = -radius; x <= radius; x += 1.0f) {
                    float pt = sqrtf(radQ - x * x);
                    for (float y = -pt; y <= pt; y += 1.0f) {
                        float blurFactor = expf(-(x * x + y * y) * invSigmaQx2) * invSigmaQx2PI;
                        float walk[4];
                        texture_linear(tex, w, h, uvx + x / (float)w, uvy + y / (float)h, walk);
                        float dC[4] = {walk[0] - centr[0], walk[1] - centr[1], walk[2] - centr[2], walk[3] - centr[3]};
                        float dd = (dC[0] * dC[0] + dC[1] * dC[1]) + (dC[2] * dC[2] + dC[3] * dC[3]);   /* glm::dot(vec4, vec4) */
                        float deltaFactor = expf(-dd * invThresholdSqx2) * invThresholdSqrt2PI * blurFactor;
                        zBuff += deltaFactor;
                        for (int c = 0; c < 4; ++c) aBuff[c] += deltaFactor * walk[c];
                    }
                }
                for (int c = 0; c < 4; ++c) col[c] = mix * (aBuff[c] / zBuff) + (1.0f - mix) * centr[c];
            }
            uint8_t* o = out + 4 * ((size_t)py * w + px);
            for (int c = 0; c < 3; ++c) {
                float v = gamma > 0.0f ? powf(col[c], 1.0f / gamma) : col[c];
                o[c] = unorm8(v);
            }
            o[3] = 255;
        }
    return VCRT_OK;
}
