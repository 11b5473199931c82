/*
 * vcrt_scene.h -- C ABI of the host-side scene producers that sit immediately before the hot path
 * (SURVEY.md 8f rows 1-2): the BVH builder of src/ray-tracing/Bvh.h, the default-scene assembly of
 * src/ray-tracing/RtScene.h and the OBJ ingestion of src/scene/mesh.cpp, plus the seeded synthetic
 * scenes BASELINE.json's configs 3-5 are measured on.  Pure host C++ behind plain C entry points;
 * outputs are arrays in the data ABI of vcrt.h, ready for vcrt_set_buffer().
 */
#ifndef VCRT_SCENE_H
#define VCRT_SCENE_H

#include "vcrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Bvh::createBvh (Bvh.h:141-209): top-down median split on a random axis (rand() % 3 per popped node, glibc's
 * default-seeded sequence reproduced internally), objects sorted by padded-box minimum (Bvh.h:116-137, pad 1e-4
 * :16,:94-98), one triangle per leaf, nodes numbered in the reference's stack order.  Bit-identical to the
 * reference's output; O(N log^2 N), parallel over subtrees.  `nodes` must hold 2*n-1 records (n > 0).
 * axis_seed: 0 = the reference's unseeded rand() sequence; otherwise srand(axis_seed). */
int vcrt_scene_build_bvh(const vcrt_triangle* triangles, uint32_t n, uint32_t axis_seed, vcrt_bvh_node* nodes, uint32_t* num_nodes);

/* Light list of RtScene.h:87-96: every triangle whose material is a LightSource, in triangle order;
 * area = 0.5*|v0 x v1| exactly as the reference computes it (position vectors, sic).  Returns the count;
 * `lights` may be NULL to query it. */
uint32_t vcrt_scene_collect_lights(const vcrt_triangle* triangles, uint32_t n, const vcrt_material* materials, uint32_t num_materials, vcrt_light* lights);

/* Seeded synthetic scene for BASELINE configs 3-5: a closed, lit box (the bundled scene's extent and camera) holding
 * a displaced-grid terrain and displaced tessellated spheres; Lambertian albedos U[0.2,0.9], one emissive ceiling quad.
 * Writes at most max_triangles triangles / max_materials materials; returns the triangle count (<= target). */
uint32_t vcrt_scene_generate_box(uint32_t target_triangles, uint32_t seed, vcrt_triangle* triangles, uint32_t max_triangles,
                                 vcrt_material* materials, uint32_t max_materials, uint32_t* num_materials);

/* OBJ ingestion as the reference's path tracer sees it (Mesh::Mesh, mesh.cpp:96-139 through tinyobjloader, then
 * getTriangles, RtScene.h:13-30): positions of the face corners, faces in file order, polygons fanned around their first
 * corner, everything else in the file ignored; every triangle gets `material_index`.  Returns the triangle count (0 with
 * vcrt_scene_last_error() set on failure); `out` may be NULL to query the count. */
uint32_t vcrt_scene_load_obj(const char* path, uint32_t material_index, vcrt_triangle* out, uint32_t max_triangles);

/* The six materials of RtScene.h:48-60 (gray, red, green, white light, metal, glass).  Returns 6. */
uint32_t vcrt_scene_default_materials(vcrt_material* out, uint32_t max_materials);

/* glibc rand() (TYPE_3 additive feedback), for tests: writes n outputs of the sequence after srand(seed). */
void vcrt_scene_glibc_rand(uint32_t seed, uint32_t n, int32_t* out);

const char* vcrt_scene_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
