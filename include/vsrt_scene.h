/*
 * vsrt_scene.h -- synthetic scene + Mesa-anv/GEN_RT_BVH arena writer and ray generators.
 *
 * Test/bench tooling, CPU only.  It has no counterpart inside the reference tree: there the BVH bytes are
 * produced by the external mesa-vulkan-sim driver + Embree (README.md:19,36-46) and handed over through
 * gpgpusim_allocTLAS/BLAS.  This writer emits the same wire format (vulkan_acceleration_structure_util.h:89-497,
 * SURVEY.md A.1) so that the reference code, the oracle and the CUDA path all consume identical bytes, and it
 * enforces the invariants the reference asserts on (vulkan_ray_tracing.cc:926,1820,2108).
 */
#ifndef VSRT_SCENE_H
#define VSRT_SCENE_H
#include <stdint.h>
#include "vsrt.h"
#ifdef __cplusplus
extern "C" {
#endif

enum { VSRT_SCENE_UNIFORM = 0, VSRT_SCENE_CLUSTERED = 1 };
#define VSRT_SCENE_F_HOLES 0x1u        /* leave empty child slots inside internal nodes */
#define VSRT_SCENE_F_TRANSFORMS 0x2u   /* random rotation/uniform scale/translation per instance (else identity) */
#define VSRT_SCENE_F_PROCEDURAL 0x4u   /* a few procedural leaves (for validators; traversal support is "next") */

typedef struct vsrt_scene_desc {
  uint64_t seed;
  uint64_t n_triangles;     /* total over all BLASes */
  uint32_t n_blas;          /* >= 1 */
  uint32_t n_instances;     /* >= n_blas; instance i references BLAS i % n_blas */
  uint32_t kind;            /* VSRT_SCENE_* */
  uint32_t flags;           /* VSRT_SCENE_F_* */
  uint32_t max_fanout;      /* 2..6, children per internal node (6 = reference-like) */
  uint32_t reserved;
} vsrt_scene_desc;

typedef struct vsrt_scene vsrt_scene;

int vsrt_scene_build(const vsrt_scene_desc* desc, vsrt_scene** out);
void vsrt_scene_free(vsrt_scene* s);
/* One contiguous 64-byte-aligned arena: TLAS header at offset 0, then TLAS nodes, then each BLAS. */
const uint8_t* vsrt_scene_arena(const vsrt_scene* s, uint64_t* size);
uint32_t vsrt_scene_n_blas(const vsrt_scene* s);
/* byte offset of BLAS b's GEN_RT_BVH header in the arena, and the size of the buffer starting there */
uint64_t vsrt_scene_blas_offset(const vsrt_scene* s, uint32_t b, uint64_t* size);
uint64_t vsrt_scene_n_nodes(const vsrt_scene* s, uint64_t* n_internal, uint64_t* n_leaves, uint32_t* depth);
/* triangle soup in object space: 9 floats per triangle, primitive index order per BLAS (for ray generation) */
const float* vsrt_scene_triangles(const vsrt_scene* s, uint64_t* n);

/* Check every invariant the reference asserts on while walking an arena from `tlas_offset`; returns 0 or
 * VSRT_E_BAD_BVH and writes a message. */
int vsrt_arena_validate(const uint8_t* arena, uint64_t size, uint64_t tlas_offset, char* msg, uint32_t msg_cap);

/* Pinhole camera at (0,0,3.5) looking down -z, vfov 45 deg.  Ray id = sample*(W*H) + y*W + x; sample 0 goes through
 * the pixel centres, the others are jittered (hash of the ray id).  Writes rays [first, first+count). */
void vsrt_rays_primary(uint32_t width, uint32_t height, uint32_t spp, uint64_t seed, uint32_t ray_flags,
                       uint64_t first, uint64_t count, vsrt_ray* out);
/* The same rays with the ids walking the frame in tile_w x tile_h pixel tiles (row-major tiles, row-major inside a tile): the
 * order of the reference's raygen launch, whose one-warp CTAs cover 8 x 4 pixels (warp_pixel_mapping WARP_8X4,
 * vulkan_ray_tracing.cc:3505).  0 x 0, or a tile that does not divide the frame, = scanline order. */
void vsrt_rays_primary_tiled(uint32_t width, uint32_t height, uint32_t spp, uint64_t seed, uint32_t ray_flags,
                             uint64_t first, uint64_t count, uint32_t tile_w, uint32_t tile_h, vsrt_ray* out);
/* Diffuse bounce: for every ray i that hit, a cosine-weighted direction about the geometric normal of the hit
 * point (flipped towards the incoming ray); misses are dropped.  Returns the number of rays written. */
uint64_t vsrt_rays_bounce(const vsrt_ray* rays, const vsrt_hit* hits, uint64_t n, uint64_t seed, uint32_t bounce,
                          uint32_t ray_flags, vsrt_ray* out);
/* Uniformly random incoherent rays inside the [-1,1]^3 scene volume. */
void vsrt_rays_random(uint64_t seed, uint32_t ray_flags, uint64_t first, uint64_t count, vsrt_ray* out);

#ifdef __cplusplus
}
#endif
#endif
