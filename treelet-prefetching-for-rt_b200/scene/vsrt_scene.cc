// vsrt_scene.cc -- synthetic triangle soup -> 6-wide quantised BVH in the Mesa-anv / GEN_RT_BVH wire format,
// plus ray generators.  CPU-only tooling (see include/vsrt_scene.h).  Wire format follows SURVEY.md A.1 /
// vulkan_acceleration_structure_util.h:89-497 of the reference; the code is written from that description.
#include "vsrt_scene.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------- counter-based RNG
inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline float u01(uint64_t key, uint64_t idx, uint32_t dim) {
  uint64_t h = splitmix(splitmix(key ^ (idx * 0xD1342543DE82EF95ull)) + dim);
  return (float)(h >> 40) * (1.0f / 16777216.0f);   // [0,1)
}

struct Box { float lo[3], hi[3]; };
inline Box empty_box() { Box b; for (int a = 0; a < 3; a++) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; } return b; }
inline void grow(Box& b, const Box& o) { for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); } }
inline void grow_pt(Box& b, const float* p) { for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], p[a]); b.hi[a] = std::max(b.hi[a], p[a]); } }

inline uint64_t expand21(uint64_t v) {   // spread 21 bits to every third bit
  v &= 0x1fffff; v = (v | v << 32) & 0x1f00000000ffffull; v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full; v = (v | v << 4) & 0x10c30c30c30c30c3ull; v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}
inline uint64_t morton63(const float c[3], const Box& b) {
  uint64_t q[3];
  for (int a = 0; a < 3; a++) {
    float e = b.hi[a] - b.lo[a];
    float t = e > 0 ? (c[a] - b.lo[a]) / e : 0.5f;
    t = std::min(std::max(t, 0.0f), 1.0f);
    q[a] = (uint64_t)std::min(2097151.0, (double)t * 2097152.0);
  }
  return (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
}

struct Prim { uint64_t code; uint32_t id; };

// ---------------------------------------------------------------- wire-format writers
inline void put_f(uint8_t* p, float f) { memcpy(p, &f, 4); }
inline void put_u32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
inline void put_u64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

void write_header(uint8_t* p, uint64_t root_offset, const Box& b) {
  memset(p, 0, 64);
  put_u64(p, root_offset);
  for (int a = 0; a < 3; a++) { put_f(p + 8 + 4 * a, b.lo[a]); put_f(p + 20 + 4 * a, b.hi[a]); }
}

// 8-bit quantisation of child boxes relative to the node origin, conservative under the decoder's own
// fp32 arithmetic  lo = origin + ldexpf(q, exp - 8)  (reference util.h:499-508).
void write_internal(uint8_t* p, int64_t child_offset_units, const uint8_t info[6], const Box child[6]) {
  memset(p, 0, 64);
  Box u = empty_box();
  for (int i = 0; i < 6; i++) if (info[i] & 3) grow(u, child[i]);
  float org[3]; int ex[3];
  for (int a = 0; a < 3; a++) {
    org[a] = u.lo[a];
    float ext = u.hi[a] - org[a];
    int e = -100;
    if (ext > 0) { int fe; frexpf(ext / 255.0f, &fe); e = fe + 8; }   // 2^(e-8) >= ext/255
    if (e < -120) e = -120;
    for (;; e++) {
      bool ok = true;
      for (int i = 0; i < 6 && ok; i++) {
        if (!(info[i] & 3)) continue;
        float s = ldexpf(1.0f, e - 8);
        float ql = floorf((child[i].lo[a] - org[a]) / s), qh = ceilf((child[i].hi[a] - org[a]) / s);
        if (ql < 0) ql = 0;
        while (ql > 0 && org[a] + ldexpf(ql, e - 8) > child[i].lo[a]) ql -= 1;
        while (qh <= 255 && org[a] + ldexpf(qh, e - 8) < child[i].hi[a]) qh += 1;
        if (qh > 255 || ql > 255) ok = false;
      }
      if (ok) break;
    }
    ex[a] = e;
  }
  for (int a = 0; a < 3; a++) put_f(p + 4 * a, org[a]);
  put_u32(p + 12, (uint32_t)(int32_t)child_offset_units);
  p[16] = 0; p[17] = 0;
  for (int a = 0; a < 3; a++) p[18 + a] = (uint8_t)(int8_t)ex[a];
  p[21] = 0xff;
  for (int i = 0; i < 6; i++) {
    p[22 + i] = info[i];
    for (int a = 0; a < 3; a++) {
      uint8_t ql = 0, qh = 0;
      if (info[i] & 3) {
        float s = ldexpf(1.0f, ex[a] - 8);
        float l = floorf((child[i].lo[a] - org[a]) / s), h = ceilf((child[i].hi[a] - org[a]) / s);
        if (l < 0) l = 0;
        while (l > 0 && org[a] + ldexpf(l, ex[a] - 8) > child[i].lo[a]) l -= 1;
        while (h < 255 && org[a] + ldexpf(h, ex[a] - 8) < child[i].hi[a]) h += 1;
        ql = (uint8_t)l; qh = (uint8_t)h;
      } else { ql = 0x80; qh = 0; }   // empty slot: inverted box, like the hardware format
      p[28 + 12 * a + i] = ql; p[28 + 12 * a + 6 + i] = qh;
    }
  }
}

void write_quad(uint8_t* p, uint32_t prim, uint32_t geom, const float* tri9) {
  memset(p, 0, 64);
  put_u32(p, 0xff000000u);                       // shader index 0, ray mask 0xff
  put_u32(p + 4, (geom & 0x0fffffffu) | (1u << 30)); // leaf type 0 (quad), GeometryFlags = opaque
  put_u32(p + 8, prim);
  put_u32(p + 12, (0u) | (0u << 17) | (1u << 19) | (2u << 21));   // PrimitiveIndex1Delta = 0 (reference assert :2108)
  for (int i = 0; i < 9; i++) put_f(p + 16 + 4 * i, tri9[i]);
  for (int i = 0; i < 3; i++) put_f(p + 52 + 4 * i, tri9[6 + i]);  // 4th vertex repeats the 3rd
}
void write_procedural(uint8_t* p, uint32_t prim, uint32_t geom) {
  memset(p, 0, 64);
  put_u32(p, 0xff000000u);
  put_u32(p + 4, (geom & 0x0fffffffu) | (1u << 29));
  put_u32(p + 8, 1u | (1u << 19));
  put_u32(p + 12, prim);
}
// row-vector 3x4: p' = p * M (3x3) + t.  Wire placement per SURVEY A.1 "matrix trap": the reference's W2O
// 4x4 is A[0..8] + B[9..11], its O2W is B[0..8] + A[9..11].
struct Xform { double m[3][3]; double t[3]; };
void write_instance(uint8_t* p, int64_t bvh_address, uint32_t instance_id, uint32_t hit_group, const Xform& w2o, const Xform& o2w) {
  memset(p, 0, 128);
  put_u32(p, 0xff000000u);
  put_u32(p + 4, (hit_group & 0x00ffffffu) | ((uint32_t)((1u /*LeafType*/ | (1u << 1)) << 5) << 24));
  p[14] = 0;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { put_f(p + 16 + 4 * (3 * r + c), (float)w2o.m[r][c]); put_f(p + 80 + 4 * (3 * r + c), (float)o2w.m[r][c]); }
  for (int c = 0; c < 3; c++) { put_f(p + 16 + 4 * (9 + c), (float)o2w.t[c]); put_f(p + 80 + 4 * (9 + c), (float)w2o.t[c]); }
  put_u64(p + 64, (uint64_t)bvh_address);
  put_u32(p + 72, instance_id);
  put_u32(p + 76, instance_id);
}

// ---------------------------------------------------------------- 6-wide builder over Morton-sorted prims
struct Builder {
  std::vector<uint8_t>& arena;
  const std::vector<Prim>& prims;     // sorted by code
  const std::vector<Box>& boxes;      // by prim id
  uint32_t fanout; bool holes; uint64_t seed;
  bool tlas;                          // leaves are instance leaves (128 B) instead of quads
  uint64_t n_internal = 0, n_leaves = 0; uint32_t max_depth = 0;
  // leaf writer: writes the leaf record for prim id at arena offset
  void (*leaf_fn)(void* user, uint64_t off, uint32_t prim_id); void* user;

  uint64_t alloc(uint64_t bytes) { uint64_t o = arena.size(); arena.resize(o + bytes); return o; }

  static uint32_t split_pos(const std::vector<Prim>& p, uint32_t lo, uint32_t hi) {
    uint64_t a = p[lo].code, b = p[hi - 1].code;
    if (a == b) return (lo + hi) / 2;
    int bit = 63 - __builtin_clzll(a ^ b);
    uint64_t mask = ~((1ull << bit) - 1);
    uint64_t pivot = (b & mask);          // first code with the differing bit set
    uint32_t l = lo, h = hi;
    while (l < h) { uint32_t m = (l + h) / 2; if (p[m].code < pivot) l = m + 1; else h = m; }
    if (l <= lo || l >= hi) return (lo + hi) / 2;
    return l;
  }

  // builds the node at arena offset node_off over [lo,hi); returns its box
  Box build(uint64_t node_off, uint32_t lo, uint32_t hi, uint32_t depth) {
    n_internal++; if (depth > max_depth) max_depth = depth;
    uint32_t rl[6], rh[6]; uint32_t nr = 1; rl[0] = lo; rh[0] = hi;
    uint32_t want = fanout;
    if (holes) { uint32_t r = (uint32_t)(splitmix(seed ^ node_off) % 4); if (want > 2 + r) want = std::max(2u, want - r); }
    while (nr < want) {
      int best = -1; uint32_t bs = 1;
      for (uint32_t i = 0; i < nr; i++) if (rh[i] - rl[i] > bs) { bs = rh[i] - rl[i]; best = (int)i; }
      if (best < 0) break;
      uint32_t m = split_pos(prims, rl[best], rh[best]);
      for (uint32_t i = nr; i > (uint32_t)best + 1; i--) { rl[i] = rl[i - 1]; rh[i] = rh[i - 1]; }
      rh[best + 1] = rh[best]; rl[best + 1] = m; rh[best] = m; nr++;
    }
    // slot assignment (optionally with empty slots in between)
    int slot_of[6]; uint32_t ns = nr;
    for (uint32_t i = 0; i < nr; i++) slot_of[i] = (int)i;
    if (holes && nr < 6) {
      uint64_t h = splitmix(seed * 31 + node_off);
      uint32_t extra = (uint32_t)(h % (6 - nr + 1));
      // choose positions: spread children over nr+extra slots
      ns = nr + extra; uint32_t k = 0;
      for (uint32_t s = 0; s < ns && k < nr; s++) {
        uint32_t remaining_slots = ns - s, remaining_children = nr - k;
        bool take = remaining_slots == remaining_children || ((splitmix(h + s) & 1) != 0);
        if (take) slot_of[k++] = (int)s;
      }
    }
    uint8_t info[6] = { 0, 0, 0, 0, 0, 0 }; Box cb[6]; uint64_t child_off[6];
    const uint32_t leaf_units = tlas ? 2 : 1;
    uint64_t block_bytes = 0;
    for (uint32_t i = 0; i < nr; i++) block_bytes += (rh[i] - rl[i] == 1 ? leaf_units : 1) * 64ull;
    uint64_t block = alloc(block_bytes), cur = block;
    for (uint32_t i = 0; i < nr; i++) {
      bool leaf = rh[i] - rl[i] == 1;
      child_off[i] = cur; cur += (leaf ? leaf_units : 1) * 64ull;
      uint32_t type = leaf ? (tlas ? 1u : 4u) : 0u;
      info[slot_of[i]] = (uint8_t)((leaf ? leaf_units : 1u) | (type << 2));
    }
    for (uint32_t i = 0; i < nr; i++) {
      if (rh[i] - rl[i] == 1) {
        uint32_t id = prims[rl[i]].id; n_leaves++;
        if (depth + 1 > max_depth) max_depth = depth + 1;
        leaf_fn(user, child_off[i], id);
        cb[slot_of[i]] = boxes[id];
      } else cb[slot_of[i]] = build(child_off[i], rl[i], rh[i], depth + 1);
    }
    // procedural marker: the leaf writer may have turned a quad into a procedural leaf -> child type 3
    if (!tlas) for (uint32_t i = 0; i < nr; i++) if (rh[i] - rl[i] == 1) {
      uint32_t dw; memcpy(&dw, &arena[child_off[i] + 4], 4);
      if ((dw >> 29) & 1) info[slot_of[i]] = (uint8_t)(1u | (3u << 2));
    }
    write_internal(&arena[node_off], ((int64_t)block - (int64_t)node_off) / 64, info, cb);
    Box u = empty_box(); for (int s = 0; s < 6; s++) if (info[s] & 3) grow(u, cb[s]);
    return u;
  }
};

}  // namespace

struct vsrt_scene {
  std::vector<uint8_t> arena_store; uint8_t* arena; uint64_t arena_size;
  std::vector<uint64_t> blas_off, blas_size, blas_tri_first;
  std::vector<float> tris;
  uint64_t n_internal, n_leaves; uint32_t depth;
};

namespace {
struct QuadCtx { std::vector<uint8_t>* arena; const float* tris; uint32_t geom; bool procedural; uint64_t seed; };
void quad_leaf_fn(void* u, uint64_t off, uint32_t id) {
  QuadCtx* c = (QuadCtx*)u;
  if (c->procedural && splitmix(c->seed ^ (0xABCDull + id)) % 97 == 0) write_procedural(&(*c->arena)[off], id, c->geom);
  else write_quad(&(*c->arena)[off], id, c->geom, c->tris + 9ull * id);
}
struct InstCtx { std::vector<uint8_t>* arena; const std::vector<uint64_t>* blas_off; const std::vector<Xform>* w2o; const std::vector<Xform>* o2w; uint32_t n_blas; };
void inst_leaf_fn(void* u, uint64_t off, uint32_t id) {
  InstCtx* c = (InstCtx*)u;
  uint64_t hdr = (*c->blas_off)[id % c->n_blas];
  write_instance(&(*c->arena)[off], (int64_t)hdr - (int64_t)off, id, id * 2 + 1, (*c->w2o)[id], (*c->o2w)[id]);
}
void invert(const Xform& a, Xform& inv) {   // p' = p*M + t  ->  p = (p' - t) * M^-1
  const double (*m)[3] = a.m;
  double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  double id = 1.0 / det;
  inv.m[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id; inv.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; inv.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  inv.m[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id; inv.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; inv.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  inv.m[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id; inv.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; inv.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  for (int c = 0; c < 3; c++) inv.t[c] = -(a.t[0] * inv.m[0][c] + a.t[1] * inv.m[1][c] + a.t[2] * inv.m[2][c]);
}
}  // namespace

extern "C" int vsrt_scene_build(const vsrt_scene_desc* d, vsrt_scene** out) {
  if (!d || !out || d->n_blas == 0 || d->n_triangles < d->n_blas) return VSRT_E_INVALID;
  const uint32_t n_blas = d->n_blas, n_inst = std::max(d->n_instances, d->n_blas);
  const uint32_t fanout = d->max_fanout >= 2 && d->max_fanout <= 6 ? d->max_fanout : 6;
  const uint64_t N = d->n_triangles;
  vsrt_scene* s = new vsrt_scene();
  s->tris.resize(9 * N);
  // ---- triangle soup: centroid U([-1,1]^3) or 64 Gaussian clusters; edge ~U(.5,1.5)*(8/N)^(1/3)
  const float base_edge = (float)cbrt(8.0 / (double)N);
  float ccen[64][3];
  for (int k = 0; k < 64; k++) for (int a = 0; a < 3; a++) ccen[k][a] = u01(d->seed ^ 0xC1u, k, a) * 1.6f - 0.8f;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)N; i++) {
    float c[3];
    if (d->kind == VSRT_SCENE_CLUSTERED) {
      int k = (int)(u01(d->seed, i, 20) * 64.0f) & 63;
      for (int a = 0; a < 3; a++) {   // Box-Muller, sigma = 0.05
        float u1 = std::max(u01(d->seed, i, 21 + 2 * a), 1e-7f), u2 = u01(d->seed, i, 22 + 2 * a);
        c[a] = ccen[k][a] + 0.05f * sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
        c[a] = std::min(std::max(c[a], -1.0f), 1.0f);
      }
    } else for (int a = 0; a < 3; a++) c[a] = u01(d->seed, i, a) * 2.0f - 1.0f;
    float edge = (0.5f + u01(d->seed, i, 3)) * base_edge;
    for (int v = 0; v < 3; v++) {      // three random points on a sphere of radius ~edge/2 around the centroid
      float z = u01(d->seed, i, 4 + 2 * v) * 2.0f - 1.0f, ph = u01(d->seed, i, 5 + 2 * v) * 6.2831853f;
      float r = sqrtf(std::max(0.0f, 1.0f - z * z));
      s->tris[9 * i + 3 * v + 0] = c[0] + 0.5f * edge * r * cosf(ph);
      s->tris[9 * i + 3 * v + 1] = c[1] + 0.5f * edge * r * sinf(ph);
      s->tris[9 * i + 3 * v + 2] = c[2] + 0.5f * edge * z;
    }
  }
  // ---- instance transforms
  std::vector<Xform> o2w(n_inst), w2o(n_inst);
  for (uint32_t i = 0; i < n_inst; i++) {
    Xform x; memset(&x, 0, sizeof(x)); x.m[0][0] = x.m[1][1] = x.m[2][2] = 1.0;
    if ((d->flags & VSRT_SCENE_F_TRANSFORMS) && i > 0) {
      double ax = u01(d->seed ^ 0x77, i, 0) * 6.2831853, ay = u01(d->seed ^ 0x77, i, 1) * 6.2831853;
      double sc = 0.5 + 1.5 * u01(d->seed ^ 0x77, i, 2);
      double cx = cos(ax), sx = sin(ax), cy = cos(ay), sy = sin(ay);
      double rx[3][3] = { { 1, 0, 0 }, { 0, cx, sx }, { 0, -sx, cx } }, ry[3][3] = { { cy, 0, -sy }, { 0, 1, 0 }, { sy, 0, cy } };
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double v = 0; for (int k = 0; k < 3; k++) v += rx[r][k] * ry[k][c]; x.m[r][c] = v * sc; }
      for (int c = 0; c < 3; c++) x.t[c] = u01(d->seed ^ 0x77, i, 3 + c) * 1.0 - 0.5;
    }
    o2w[i] = x; invert(x, w2o[i]);
    // round-trip through float so both matrices are the exact wire values
  }
  // ---- arena: [TLAS header][TLAS root + nodes + instance leaves] [BLAS 0] [BLAS 1] ...
  std::vector<uint8_t>& A = s->arena_store;
  A.reserve((size_t)(N * 64 * 3 / 2 + 65536));
  A.resize(64 + 64);                       // TLAS header @0, TLAS root internal @64
  // BLASes first need their offsets for the instance leaves -> build BLASes after reserving the TLAS region.
  // TLAS size is bounded: internal nodes <= n_inst, leaves n_inst*128.  Reserve exactly by building TLAS later
  // into its own vector and BLASes with offsets relative to a known TLAS size: do two-phase: build TLAS
  // topology size first via a dry run with dummy offsets.
  std::vector<Box> blas_box(n_blas);
  std::vector<std::vector<uint8_t>> blas_arena(n_blas);
  s->blas_tri_first.resize(n_blas + 1);
  uint64_t n_int = 0, n_leaf = 0; uint32_t depth = 0;
  for (uint32_t b = 0; b < n_blas; b++) {
    uint64_t t0 = N * b / n_blas, t1 = N * (b + 1) / n_blas; s->blas_tri_first[b] = t0;
    uint32_t n = (uint32_t)(t1 - t0);
    const float* tri = &s->tris[9 * t0];
    std::vector<Box> boxes(n); std::vector<Prim> prims(n);
    Box cb = empty_box();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) { Box bx = empty_box(); for (int v = 0; v < 3; v++) grow_pt(bx, tri + 9 * i + 3 * v); boxes[i] = bx; }
    for (uint32_t i = 0; i < n; i++) { float c[3]; for (int a = 0; a < 3; a++) c[a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]); grow_pt(cb, c); }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) { float c[3]; for (int a = 0; a < 3; a++) c[a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]); prims[i].code = morton63(c, cb); prims[i].id = (uint32_t)i; }
    std::sort(prims.begin(), prims.end(), [](const Prim& x, const Prim& y) { return x.code < y.code || (x.code == y.code && x.id < y.id); });
    std::vector<uint8_t>& BA = blas_arena[b];
    BA.reserve((size_t)n * 88 + 4096);
    BA.resize(128);                        // header @0, root internal @64
    QuadCtx qc = { &BA, tri, b, (d->flags & VSRT_SCENE_F_PROCEDURAL) != 0, d->seed };
    Builder bl = { BA, prims, boxes, fanout, (d->flags & VSRT_SCENE_F_HOLES) != 0, d->seed + b, false };
    bl.leaf_fn = quad_leaf_fn; bl.user = &qc;
    Box rb;
    if (n == 1) {  // degenerate: root with a single leaf child
      rb = bl.build(64, 0, 1, 1);
    } else rb = bl.build(64, 0, n, 1);
    write_header(&BA[0], 64, rb);
    blas_box[b] = rb; n_int += bl.n_internal; n_leaf += bl.n_leaves; depth = std::max(depth, bl.max_depth);
  }
  s->blas_tri_first[n_blas] = N;
  // TLAS over instance world boxes
  std::vector<Box> ibox(n_inst); std::vector<Prim> iprims(n_inst);
  Box icb = empty_box();
  for (uint32_t i = 0; i < n_inst; i++) {
    const Box& ob = blas_box[i % n_blas]; Box wb = empty_box();
    for (int k = 0; k < 8; k++) {
      float p[3] = { (k & 1) ? ob.hi[0] : ob.lo[0], (k & 2) ? ob.hi[1] : ob.lo[1], (k & 4) ? ob.hi[2] : ob.lo[2] }, q[3];
      for (int c = 0; c < 3; c++) q[c] = (float)(p[0] * o2w[i].m[0][c] + p[1] * o2w[i].m[1][c] + p[2] * o2w[i].m[2][c] + o2w[i].t[c]);
      grow_pt(wb, q);
    }
    for (int a = 0; a < 3; a++) { float pad = 1e-4f * (1.0f + fabsf(wb.lo[a]) + fabsf(wb.hi[a])); wb.lo[a] -= pad; wb.hi[a] += pad; }
    ibox[i] = wb; float c[3]; for (int a = 0; a < 3; a++) c[a] = 0.5f * (wb.lo[a] + wb.hi[a]); grow_pt(icb, c);
  }
  for (uint32_t i = 0; i < n_inst; i++) { float c[3]; for (int a = 0; a < 3; a++) c[a] = 0.5f * (ibox[i].lo[a] + ibox[i].hi[a]); iprims[i].code = morton63(c, icb); iprims[i].id = i; }
  std::sort(iprims.begin(), iprims.end(), [](const Prim& x, const Prim& y) { return x.code < y.code || (x.code == y.code && x.id < y.id); });
  // dry run for the TLAS size (offsets of BLAS headers depend on it)
  s->blas_off.assign(n_blas, 0); s->blas_size.assign(n_blas, 0);
  uint64_t tlas_bytes = 0;
  for (int pass = 0; pass < 2; pass++) {
    A.resize(128);
    InstCtx ic = { &A, &s->blas_off, &w2o, &o2w, n_blas };
    Builder tb = { A, iprims, ibox, fanout, (d->flags & VSRT_SCENE_F_HOLES) != 0, d->seed ^ 0x71A5, true };
    tb.leaf_fn = inst_leaf_fn; tb.user = &ic;
    Box rb = tb.build(64, 0, n_inst, 1);
    write_header(&A[0], 64, rb);
    if (pass == 0) {
      tlas_bytes = A.size();
      uint64_t cur = tlas_bytes;
      for (uint32_t b = 0; b < n_blas; b++) { s->blas_off[b] = cur; s->blas_size[b] = blas_arena[b].size(); cur += blas_arena[b].size(); }
    } else { n_int += tb.n_internal; n_leaf += tb.n_leaves; depth += tb.max_depth; }
  }
  for (uint32_t b = 0; b < n_blas; b++) { A.insert(A.end(), blas_arena[b].begin(), blas_arena[b].end()); std::vector<uint8_t>().swap(blas_arena[b]); }
  // 64-byte aligned view
  A.resize(A.size() + 64);
  uintptr_t p = (uintptr_t)A.data(); uintptr_t ap = (p + 63) & ~(uintptr_t)63;
  if (ap != p) memmove((void*)ap, (void*)p, A.size() - 64);
  s->arena = (uint8_t*)ap; s->arena_size = A.size() - 64;
  s->n_internal = n_int; s->n_leaves = n_leaf; s->depth = depth;
  *out = s;
  return VSRT_OK;
}

extern "C" void vsrt_scene_free(vsrt_scene* s) { delete s; }
extern "C" const uint8_t* vsrt_scene_arena(const vsrt_scene* s, uint64_t* size) { if (size) *size = s->arena_size; return s->arena; }
extern "C" uint32_t vsrt_scene_n_blas(const vsrt_scene* s) { return (uint32_t)s->blas_off.size(); }
extern "C" uint64_t vsrt_scene_blas_offset(const vsrt_scene* s, uint32_t b, uint64_t* size) { if (size) *size = s->blas_size[b]; return s->blas_off[b]; }
extern "C" uint64_t vsrt_scene_n_nodes(const vsrt_scene* s, uint64_t* ni, uint64_t* nl, uint32_t* depth) {
  if (ni) *ni = s->n_internal; if (nl) *nl = s->n_leaves; if (depth) *depth = s->depth; return s->n_internal + s->n_leaves;
}
extern "C" const float* vsrt_scene_triangles(const vsrt_scene* s, uint64_t* n) { if (n) *n = s->tris.size() / 9; return s->tris.data(); }

// ---------------------------------------------------------------- validator
namespace {
struct Val { const uint8_t* a; uint64_t size; char* msg; uint32_t cap; uint64_t visited; };
bool fail(Val& v, const char* what, uint64_t off) { if (v.msg && v.cap) snprintf(v.msg, v.cap, "%s at arena offset %llu", what, (unsigned long long)off); return false; }
bool walk(Val& v, uint64_t off, bool top, uint32_t depth) {
  if (depth > 200) return fail(v, "tree deeper than 200 levels", off);
  if (off % 64 || off + 64 > v.size) return fail(v, "internal node out of bounds / unaligned", off);
  const uint8_t* p = v.a + off; int32_t co; memcpy(&co, p + 12, 4);
  int64_t child = (int64_t)off + (int64_t)co * 64;
  v.visited++;
  for (int i = 0; i < 6; i++) {
    uint8_t t = p[22 + i] & 0x3f; uint32_t sz = t & 3, ty = t >> 2;
    if (sz) {
      if (child < 0 || (uint64_t)child + sz * 64ull > v.size) return fail(v, "child out of bounds", off);
      if (ty == 0) { if (sz != 1) return fail(v, "internal child with size != 1", off); if (!walk(v, (uint64_t)child, top, depth + 1)) return false; }
      else if (top) {
        if (ty != 1) return fail(v, "TLAS leaf that is not an instance (reference assert :1820)", off);
        if (sz != 2) return fail(v, "instance leaf with size != 2", off);
        uint64_t bvh; memcpy(&bvh, v.a + child + 64, 8);
        if (bvh == 0) return fail(v, "instance leaf with BVHAddress == 0 (reference assert :1900)", (uint64_t)child);
        int64_t hdr = child + (int64_t)bvh;
        if (hdr < 0 || hdr % 64 || (uint64_t)hdr + 64 > v.size) return fail(v, "BLAS header out of bounds", (uint64_t)child);
        uint64_t ro; memcpy(&ro, v.a + hdr, 8);
        if (!walk(v, (uint64_t)hdr + ro, false, depth + 1)) return false;
      } else {
        if (ty == 1) return fail(v, "instance leaf inside a BLAS", off);
        if (sz != 1) return fail(v, "BLAS leaf with size != 1", off);
        uint32_t dw, w3; memcpy(&dw, v.a + child + 4, 4); memcpy(&w3, v.a + child + 12, 4);
        if (((dw >> 29) & 1) == 0 && (w3 & 0x1ffff) != 0) return fail(v, "quad leaf with PrimitiveIndex1Delta != 0 (reference assert :2108)", (uint64_t)child);
        v.visited++;
      }
    }
    child += sz * 64;
  }
  return true;
}
}  // namespace
extern "C" int vsrt_arena_validate(const uint8_t* arena, uint64_t size, uint64_t tlas_offset, char* msg, uint32_t cap) {
  Val v = { arena, size, msg, cap, 0 };
  if (msg && cap) msg[0] = 0;
  if (tlas_offset % 64 || tlas_offset + 64 > size) { fail(v, "TLAS header out of bounds", tlas_offset); return VSRT_E_BAD_BVH; }
  uint64_t ro; memcpy(&ro, arena + tlas_offset, 8);
  return walk(v, tlas_offset + ro, true, 1) ? VSRT_OK : VSRT_E_BAD_BVH;
}

// ---------------------------------------------------------------- rays
static inline void set_ray(vsrt_ray* r, const float o[3], const float d[3], float tmin, float tmax, uint32_t flags) {
  for (int a = 0; a < 3; a++) { r->origin[a] = o[a]; r->direction[a] = d[a]; }
  r->tmin = tmin; r->tmax = tmax; r->ray_flags = flags; r->cull_mask = 0xff; r->sbt_record_offset = 0; r->sbt_record_stride = 0; r->miss_index = 0;
}
// tile_w x tile_h > 0: ray ids walk the frame in pixel tiles, one tile after the other in row-major tile order and row-major inside
// a tile -- the order in which the reference's raygen launch hands rays to traceRay: it launches CTAs of one warp that cover
// 8 x 4 pixels (warp_pixel_mapping mapping = WARP_8X4, vulkan_ray_tracing.cc:3505, :3542-3564), grid x-major.  The rays are
// the scanline generator's, permuted: jitter is keyed by the pixel's scanline id.  A frame that the tile does not divide
// falls back to scanlines.
extern "C" void vsrt_rays_primary_tiled(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint32_t flags, uint64_t first, uint64_t count,
                                        uint32_t tile_w, uint32_t tile_h, vsrt_ray* out) {
  const float tan_half = tanf(0.5f * 45.0f * 3.14159265f / 180.0f), aspect = (float)W / (float)H;
  const bool tiled = tile_w && tile_h && W % tile_w == 0 && H % tile_h == 0;
  const uint64_t frame = (uint64_t)W * H, tile_px = (uint64_t)tile_w * tile_h, tiles_x = tiled ? W / tile_w : 1;
  (void)spp;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)count; k++) {
    uint64_t id = first + (uint64_t)k;
    // ray id = sample * (W*H) + pixel: a contiguous block of W*H ids is one full frame of one sample, so ranks
    // that take consecutive blocks get statistically identical work
    const uint64_t sm = id / frame, p = id % frame;
    uint32_t x, y;
    if (tiled) {
      const uint64_t t = p / tile_px, i = p % tile_px;
      x = (uint32_t)((t % tiles_x) * tile_w + i % tile_w); y = (uint32_t)((t / tiles_x) * tile_h + i / tile_w);
    } else { x = (uint32_t)(p % W); y = (uint32_t)(p / W); }
    const uint64_t pid = sm * frame + (uint64_t)y * W + x;      // the ray's id in scanline order: keys the jitter
    float jx = 0.5f, jy = 0.5f;
    if (sm > 0) { jx = u01(seed, pid, 0); jy = u01(seed, pid, 1); }   // sample 0 = pixel centres, the others jittered
    float px = (((float)x + jx) / (float)W * 2.0f - 1.0f) * tan_half * aspect;
    float py = (1.0f - ((float)y + jy) / (float)H * 2.0f) * tan_half;
    float d[3] = { px, py, -1.0f }; float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int a = 0; a < 3; a++) d[a] /= n;
    float o[3] = { 0.0f, 0.0f, 3.5f };
    set_ray(&out[k], o, d, 1e-3f, 1e30f, flags);
  }
}
extern "C" void vsrt_rays_primary(uint32_t W, uint32_t H, uint32_t spp, uint64_t seed, uint32_t flags, uint64_t first, uint64_t count, vsrt_ray* out) {
  vsrt_rays_primary_tiled(W, H, spp, seed, flags, first, count, 0, 0, out);
}
extern "C" void vsrt_rays_random(uint64_t seed, uint32_t flags, uint64_t first, uint64_t count, vsrt_ray* out) {
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)count; k++) {
    uint64_t id = first + (uint64_t)k;
    float o[3], d[3];
    for (int a = 0; a < 3; a++) o[a] = u01(seed, id, a) * 2.0f - 1.0f;
    float z = u01(seed, id, 3) * 2.0f - 1.0f, ph = u01(seed, id, 4) * 6.2831853f, r = sqrtf(std::max(0.0f, 1.0f - z * z));
    d[0] = r * cosf(ph); d[1] = r * sinf(ph); d[2] = z;
    set_ray(&out[k], o, d, 1e-3f, 1e30f, flags);
  }
}
// The scene's triangles are needed for the geometric normal; the caller passes them (vsrt_scene_triangles) via
// a module-level pointer set by vsrt_rays_bounce_scene to keep the C signature plain.
static const vsrt_scene* g_bounce_scene = nullptr;
extern "C" void vsrt_rays_bounce_scene(const vsrt_scene* s) { g_bounce_scene = s; }
extern "C" uint64_t vsrt_rays_bounce(const vsrt_ray* rays, const vsrt_hit* hits, uint64_t n, uint64_t seed, uint32_t bounce, uint32_t flags, vsrt_ray* out) {
  const vsrt_scene* s = g_bounce_scene;
  uint64_t w = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (!hits[i].hit_geometry) continue;
    float nrm[3] = { 0, 0, 1 };
    if (s) {
      uint32_t g = hits[i].geometry_index; uint64_t t = (g < s->blas_tri_first.size() ? s->blas_tri_first[g] : 0) + hits[i].primitive_index;
      if (9 * t + 8 < s->tris.size()) {
        const float* p = &s->tris[9 * t];
        float e1[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] }, e2[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
        nrm[0] = e1[1] * e2[2] - e1[2] * e2[1]; nrm[1] = e1[2] * e2[0] - e1[0] * e2[2]; nrm[2] = e1[0] * e2[1] - e1[1] * e2[0];
        float l = sqrtf(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
        if (l > 0) for (int a = 0; a < 3; a++) nrm[a] /= l; else { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
      }
    }
    const float* din = rays[i].direction;
    if (nrm[0] * din[0] + nrm[1] * din[1] + nrm[2] * din[2] > 0) for (int a = 0; a < 3; a++) nrm[a] = -nrm[a];
    // cosine-weighted hemisphere about nrm
    float u1 = u01(seed + bounce * 0x9E37ull, i, 0), u2 = u01(seed + bounce * 0x9E37ull, i, 1);
    float r = sqrtf(u1), ph = 6.2831853f * u2, lx = r * cosf(ph), ly = r * sinf(ph), lz = sqrtf(std::max(0.0f, 1.0f - u1));
    float t1[3], t2[3];
    if (fabsf(nrm[0]) > 0.5f) { t1[0] = -nrm[1]; t1[1] = nrm[0]; t1[2] = 0; } else { t1[0] = 0; t1[1] = -nrm[2]; t1[2] = nrm[1]; }
    float l1 = sqrtf(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]); for (int a = 0; a < 3; a++) t1[a] /= l1;
    t2[0] = nrm[1] * t1[2] - nrm[2] * t1[1]; t2[1] = nrm[2] * t1[0] - nrm[0] * t1[2]; t2[2] = nrm[0] * t1[1] - nrm[1] * t1[0];
    float d[3], o[3];
    for (int a = 0; a < 3; a++) { d[a] = lx * t1[a] + ly * t2[a] + lz * nrm[a]; o[a] = hits[i].intersection_point[a] + 1e-3f * nrm[a]; }
    float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); for (int a = 0; a < 3; a++) d[a] /= dl;
    set_ray(&out[w++], o, d, 1e-3f, 1e30f, flags);
  }
  return w;
}
