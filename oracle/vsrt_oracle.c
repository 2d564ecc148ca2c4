/*
 * vsrt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's functional ray-traversal path
 * (Vulkan-Sim, ubc-aamodt-group/treelet-prefetching-for-rt).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product library (libvsrt.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED against the reference's own code.  The reference ships no
 * tests or golden vectors for this path, so the pin is
 * oracle/_ref/libvsrt_ref.so -- the reference's own function bodies compiled from
 * /root/reference by oracle/build_ref.sh -- on the scenes of tests/test_oracle.py and
 * the fixtures under tests/golden/ that were generated from it.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src).  The arithmetic is single-precision, no FMA contraction,
 * ternary MIN/MAX exactly as the reference macros; compile with
 *   gcc -O2 -std=c99 -ffp-contract=off -fopenmp
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- ABI structs */
typedef struct { float origin[3]; float tmin; float dir[3]; float tmax;
                 uint32_t flags, cull_mask, sbt_offset, sbt_stride, miss_index; } vo_ray;
typedef struct { uint32_t hit; float t; uint32_t prim, geom, instance_id; float bary[3]; float point[3];
                 uint32_t n_all_hits; } vo_hit;
typedef struct { uint64_t address; uint32_t size; uint32_t type; } vo_txn;
typedef struct { uint64_t mem_access_type[9];
                 uint64_t num_hits, num_any_hits, n_anyhit_rays, n_closesthit_rays;
                 uint64_t max_nodes_per_ray, tot_nodes_per_ray, max_tree_depth, accessed_data_size, ray_count; } vo_counters;

enum { T_STRUCT = 0, T_INTERNAL = 1, T_INSTANCE = 2, T_DESC = 3, T_QUAD = 4, T_QUAD_HIT = 5, T_PROC = 6 };
enum { NODE_INTERNAL = 0, NODE_INSTANCE = 1, NODE_PROCEDURAL = 3, NODE_QUAD = 4 };

/* ---------------------------------------------------------------- u64 -> u64 hash map */
typedef struct { uint64_t* k; uint64_t* v; uint64_t cap, n; } map64;
static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
static void map_init(map64* m, uint64_t cap) {
  uint64_t c = 16; while (c < cap * 2) c <<= 1;
  m->k = (uint64_t*)malloc(c * 8); m->v = (uint64_t*)malloc(c * 8); m->cap = c; m->n = 0;
  memset(m->k, 0xff, c * 8);
}
static void map_free(map64* m) { free(m->k); free(m->v); m->k = m->v = NULL; m->cap = m->n = 0; }
static void map_put(map64* m, uint64_t key, uint64_t val);
static void map_grow(map64* m) {
  map64 b; map_init(&b, m->cap);
  for (uint64_t i = 0; i < m->cap; i++) if (m->k[i] != ~0ull) map_put(&b, m->k[i], m->v[i]);
  map_free(m); *m = b;
}
static void map_put(map64* m, uint64_t key, uint64_t val) {
  if ((m->n + 1) * 2 > m->cap) map_grow(m);
  uint64_t i = mix64(key) & (m->cap - 1);
  while (m->k[i] != ~0ull && m->k[i] != key) i = (i + 1) & (m->cap - 1);
  if (m->k[i] == ~0ull) { m->k[i] = key; m->n++; }
  m->v[i] = val;
}
static int map_get(const map64* m, uint64_t key, uint64_t* val) {
  if (!m->cap) return 0;
  uint64_t i = mix64(key) & (m->cap - 1);
  while (m->k[i] != ~0ull) { if (m->k[i] == key) { *val = m->v[i]; return 1; } i = (i + 1) & (m->cap - 1); }
  return 0;
}

/* ---------------------------------------------------------------- context */
typedef struct { uint64_t root; uint64_t first; uint32_t count; } vo_treelet;   /* root = device address */
typedef struct { uint64_t addr; uint32_t size; } vo_lnode;
typedef struct vo_ctx {
  uint64_t tlas_dev;                 /* VulkanRayTracing::tlas_addr (vulkan_ray_tracing.cc:4896-4899) */
  map64 blas;                        /* blas_addr_map: host header -> device (:4891-4894) */
  int formed; int budget;
  vo_treelet* treelets; uint64_t n_treelets;   /* ascending root order */
  vo_lnode* lnodes; uint64_t n_lnodes;
  map64 node_root;                   /* node_map_addr_only (:475-489) */
  map64 root_idx;                    /* treelet_addr_to_metadata_idx (:1332) */
  uint64_t total_bvh_size;
  vo_counters c;
  uint64_t* proc_sink; uint64_t proc_cap, proc_n;   /* optional: instance leaf (host address) of every procedural-leaf visit, in trace order (single-threaded traces) */
} vo_ctx;

vo_ctx* vo_create(void) { vo_ctx* c = (vo_ctx*)calloc(1, sizeof(vo_ctx)); map_init(&c->blas, 16); return c; }
static void vo_clear_treelets(vo_ctx* c) {
  free(c->treelets); free(c->lnodes); c->treelets = NULL; c->lnodes = NULL; c->n_treelets = c->n_lnodes = 0;
  if (c->node_root.cap) map_free(&c->node_root);
  if (c->root_idx.cap) map_free(&c->root_idx);
  c->formed = 0;
}
void vo_destroy(vo_ctx* c) { if (!c) return; vo_clear_treelets(c); map_free(&c->blas); free(c); }
void vo_alloc_tlas(vo_ctx* c, const void* host, uint64_t size, uint64_t dev) { (void)host; (void)size; c->tlas_dev = dev; }
void vo_alloc_blas(vo_ctx* c, const void* host, uint64_t size, uint64_t dev) { (void)size; map_put(&c->blas, (uint64_t)(uintptr_t)host, dev); }
void vo_reset_counters(vo_ctx* c) { memset(&c->c, 0, sizeof(c->c)); }
void vo_set_proc_sink(vo_ctx* c, uint64_t* buf, uint64_t cap) { c->proc_sink = buf; c->proc_cap = cap; c->proc_n = 0; }
uint64_t vo_proc_sink_count(vo_ctx* c) { return c->proc_n; }
static void proc_visit(vo_ctx* c, const uint8_t* instance_leaf) {
  if (c->proc_sink && c->proc_n < c->proc_cap) c->proc_sink[c->proc_n] = (uint64_t)(uintptr_t)instance_leaf;
  c->proc_n++;
}
void vo_get_counters(vo_ctx* c, vo_counters* out) { *out = c->c; }

/* ---------------------------------------------------------------- wire-format decode */
static inline float ldf(const uint8_t* p) { float f; memcpy(&f, p, 4); return f; }
static inline uint32_t ldu32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ldu64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* GEN_RT_BVH_INTERNAL_NODE_unpack, vulkan_acceleration_structure_util.h:134-207 */
typedef struct { float org[3]; int32_t child_offset; int ex[3]; uint32_t size[6], type[6]; const uint8_t* q; } inode;
static void unpack_internal(inode* n, const uint8_t* d) {
  n->org[0] = ldf(d); n->org[1] = ldf(d + 4); n->org[2] = ldf(d + 8);
  n->child_offset = (int32_t)ldu32(d + 12);
  n->ex[0] = (int8_t)d[18]; n->ex[1] = (int8_t)d[19]; n->ex[2] = (int8_t)d[20];
  for (int i = 0; i < 6; i++) { uint8_t t = d[22 + i] & 0x3f; n->size[i] = t & 3; n->type[i] = t >> 2; }
  n->q = d + 28;   /* LowerX[6] UpperX[6] LowerY[6] UpperY[6] LowerZ[6] UpperZ[6] */
}
/* set_child_bounds, util.h:499-508: lo = Origin + ldexpf((float)q, exp - 8) */
static void child_bounds(const inode* n, int c, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; a++) {
    lo[a] = n->org[a] + ldexpf((float)n->q[12 * a + c], n->ex[a] - 8);
    hi[a] = n->org[a] + ldexpf((float)n->q[12 * a + 6 + c], n->ex[a] - 8);
  }
}
/* instance leaf: util.h:271-386 + instance_leaf_matrix_to_float4x4 :210-224.  The 4x4 the reference builds
 * reads 12 contiguous STRUCT floats, i.e. wire floats A[0..8] then B[9..11] (SURVEY A.1 "matrix trap"). */
typedef struct { float w2o[4][4]; uint64_t bvh_address; uint32_t instance_id; uint32_t hit_group; } ileaf;
static void unpack_instance(ileaf* l, const uint8_t* d) {
  const uint8_t* A = d + 16; const uint8_t* B = d + 80;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) l->w2o[r][c] = ldf(A + 4 * (3 * r + c)); l->w2o[r][3] = 0.0f; }
  for (int c = 0; c < 3; c++) l->w2o[3][c] = ldf(B + 4 * (9 + c));
  l->w2o[3][3] = 1.0f;
  l->bvh_address = ldu64(d + 64); l->instance_id = ldu32(d + 72);
  l->hit_group = ldu32(d + 4) & 0x00ffffffu;
}

/* ---------------------------------------------------------------- ray math */
typedef struct { float o[3], d[3], tmin, tmax; } rayf;
#define RMIN(a, b) (((a) < (b)) ? (a) : (b))   /* vulkan_ray_tracing.h:53-56, NaN -> second operand */
#define RMAX(a, b) (((a) > (b)) ? (a) : (b))

/* calculate_idir, vulkan_ray_tracing.cc:220-235 */
static void calc_idir(const float d[3], float idir[3]) {
  const float ooeps = 0x1p-80f;
  for (int a = 0; a < 3; a++) idir[a] = 1.0f / (fabsf(d[a]) > ooeps ? d[a] : copysignf(ooeps, d[a]));
}
/* ray_box_test + get_t_bound + magic_max7/min7, :183-257 */
static int ray_box(const float lo[3], const float hi[3], const float idir[3], const float o[3], float tmin, float tmax, float* thit) {
  float l[3], h[3];
  for (int a = 0; a < 3; a++) { l[a] = (lo[a] - o[a]) * idir[a]; h[a] = (hi[a] - o[a]) * idir[a]; }
  float t1 = RMAX(RMIN(l[0], h[0]), tmin); float t2 = RMAX(RMIN(l[1], h[1]), t1); float mn = RMAX(RMIN(l[2], h[2]), t2);
  float u1 = RMIN(RMAX(l[0], h[0]), tmax); float u2 = RMIN(RMAX(l[1], h[1]), u1); float mx = RMIN(RMAX(l[2], h[2]), u2);
  *thit = mn;
  return mn <= mx;
}
/* make_transformed_ray, :168-181; float4x4::operator*, util.h:47-55 (row vector, accumulate from 0) */
static void transform_ray(const rayf* r, const float m[4][4], rayf* out, float* tmult) {
  float vo[4] = { r->o[0], r->o[1], r->o[2], 1.0f }, vd[4] = { r->d[0], r->d[1], r->d[2], 0.0f };
  float ro[4], rd[4];
  for (int i = 0; i < 4; i++) {
    float so = 0.0f, sd = 0.0f;
    for (int j = 0; j < 4; j++) { so += m[j][i] * vo[j]; sd += m[j][i] * vd[j]; }
    ro[i] = so; rd[i] = sd;
  }
  out->o[0] = ro[0] / ro[3]; out->o[1] = ro[1] / ro[3]; out->o[2] = ro[2] / ro[3];
  float norm = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);   /* get_norm(float3), :152-155 */
  *tmult = norm;
  out->d[0] = rd[0] / norm; out->d[1] = rd[1] / norm; out->d[2] = rd[2] / norm;   /* normalized, :162-166 */
  out->tmin = r->tmin * norm; out->tmax = r->tmax * norm;
}
/* mt_ray_triangle_test, :3089-3111; cross/dot: gpgpu-sim/vector-math.cc:41-50 */
static void crossf(const float a[3], const float b[3], float r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
static float dotf(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static int ray_tri(const float p0[3], const float p1[3], const float p2[3], const rayf* r, float* thit) {
  float e1[3], e2[3], pv[3], tv[3], qv[3];
  for (int a = 0; a < 3; a++) { e1[a] = p1[a] - p0[a]; e2[a] = p2[a] - p0[a]; }
  crossf(r->d, e2, pv);
  float det = dotf(e1, pv);
  float idet = 1 / det;
  for (int a = 0; a < 3; a++) tv[a] = r->o[a] - p0[a];
  float u = dotf(tv, pv) * idet;
  if (u < 0 || u > 1) return 0;
  crossf(tv, e1, qv);
  float v = dotf(r->d, qv) * idet;
  if (v < 0 || (u + v) > 1) return 0;
  *thit = dotf(e2, qv) * idet;
  return 1;
}
/* Barycentric, :3113-3130 -> {v, w, u} */
static void barycentric(const float p[3], const float a[3], const float b[3], const float c[3], float out[3]) {
  float v0[3], v1[3], v2[3];
  for (int i = 0; i < 3; i++) { v0[i] = b[i] - a[i]; v1[i] = c[i] - a[i]; v2[i] = p[i] - a[i]; }
  float d00 = dotf(v0, v0), d01 = dotf(v0, v1), d11 = dotf(v1, v1), d20 = dotf(v2, v0), d21 = dotf(v2, v1);
  float denom = d00 * d11 - d01 * d01;
  float v = (d11 * d20 - d01 * d21) / denom;
  float w = (d00 * d21 - d01 * d20) / denom;
  float u = 1.0f - v - w;
  out[0] = v; out[1] = w; out[2] = u;
}

/* ---------------------------------------------------------------- treelet formation
 * createTreelets, vulkan_ray_tracing.cc:823-1470.  FIFO `stack` of candidates with byte costs, FIFO of
 * pending roots; only the FRONT candidate is tested against the remaining budget (:1117); when it does not
 * fit every queued candidate becomes a future root (:1167-1171).  A treelet is a pure function of
 * (root, budget), so a root address met twice is formed once (the reference re-forms and overwrites the
 * same map slot, :1159). */
typedef struct { const uint8_t* addr; uint8_t top, leaf; int cost; } cand;
typedef struct { cand* v; uint64_t head, tail, cap; } fifo;
static void fifo_push(fifo* f, cand c) {
  if (f->tail == f->cap) {
    if (f->head > f->cap / 2) { memmove(f->v, f->v + f->head, (f->tail - f->head) * sizeof(cand)); f->tail -= f->head; f->head = 0; }
    else { f->cap = f->cap ? f->cap * 2 : 1024; f->v = (cand*)realloc(f->v, f->cap * sizeof(cand)); }
  }
  f->v[f->tail++] = c;
}
static int cmp_treelet(const void* a, const void* b) {
  uint64_t x = ((const vo_treelet*)a)->root, y = ((const vo_treelet*)b)->root; return x < y ? -1 : x > y;
}

int vo_form_treelets(vo_ctx* c, const void* tlas_v, int budget) {
  if (c->formed) return 0;
  const uint8_t* tlas = (const uint8_t*)tlas_v;
  const int64_t off = (int64_t)(c->tlas_dev - (uint64_t)(uintptr_t)tlas);   /* device_offset, :1549 */
  fifo pend = { 0 }, st = { 0 };
  map64 done; map_init(&done, 1024);
  vo_treelet* tl = NULL; uint64_t ntl = 0, captl = 0;
  vo_lnode* ln = NULL; uint64_t nln = 0, capln = 0;
  uint64_t total_bvh = 0;
  int err = 0;
#define LPUSH(a_, s_) do { if (nln == capln) { capln = capln ? capln * 2 : 4096; ln = (vo_lnode*)realloc(ln, capln * sizeof(vo_lnode)); } \
                           ln[nln].addr = (a_); ln[nln].size = (s_); nln++; } while (0)
  /* :846-862 first treelet is keyed by the TLAS header and starts with header + root internal node */
  cand first = { tlas, 1, 0, 64 };
  fifo_push(&pend, first);
  int is_first = 1;
  while (pend.head < pend.tail && !err) {
    cand root = pend.v[pend.head++];
    uint64_t root_dev = (uint64_t)(uintptr_t)root.addr + (uint64_t)off;   /* roots are never BLAS headers, :1143 */
    uint64_t dummy;
    if (map_get(&done, root_dev, &dummy)) continue;
    map_put(&done, root_dev, 1);
    int remaining = budget;
    uint64_t list_first = nln;
    st.head = st.tail = 0;
    cand next;
    if (is_first) {
      remaining -= 64; total_bvh += 64;                                  /* :853-856 */
      if (remaining < 0) { err = -8; break; }
      LPUSH(root_dev, 64);
      next.addr = tlas + ldu64(tlas); next.top = 1; next.leaf = 0; next.cost = 64;   /* :858-861 */
      is_first = 0;
    } else next = root;                                                  /* processed without a fit check, :1197 */
    for (;;) {
      /* ---- process `next` (:886-1114) ---- */
      if (!next.leaf) {                                                  /* TLAS or BLAS internal node */
        inode n; unpack_internal(&n, next.addr);
        remaining -= 64; total_bvh += 64;
        if (remaining < 0) { err = -8; break; }
        LPUSH((uint64_t)(uintptr_t)next.addr + (uint64_t)off, 64);
        const uint8_t* child = next.addr + (int64_t)n.child_offset * 64;
        for (int i = 0; i < 6; i++) {
          if (n.size[i] > 0) {
            cand ch; ch.addr = child; ch.top = next.top;
            if (n.type[i] != NODE_INTERNAL) {
              if (next.top && n.type[i] != NODE_INSTANCE) { err = -6; break; }   /* assert :926 */
              ch.leaf = 1; ch.cost = next.top ? 128 + 64 : 64;           /* :927-928, :1059-1060 */
            } else { ch.leaf = 0; ch.cost = 64; }
            fifo_push(&st, ch);
          }
          child += n.size[i] * 64;
        }
        if (err) break;
      } else if (next.top) {                                             /* instance leaf, :944-993 */
        ileaf l; unpack_instance(&l, next.addr);
        remaining -= 128; total_bvh += 128;
        if (remaining < 0) { err = -8; break; }
        LPUSH((uint64_t)(uintptr_t)next.addr + (uint64_t)off, 128);
        const uint8_t* hdr = next.addr + l.bvh_address;
        uint64_t hdr_dev;
        if (!map_get(&c->blas, (uint64_t)(uintptr_t)hdr, &hdr_dev)) { err = -5; break; }   /* assert :973 */
        remaining -= 64; total_bvh += 64;
        LPUSH(hdr_dev, 64);                                              /* isBlasRoot entry, :975, :1149-1153 */
        if (remaining < 0) { err = -8; break; }
        cand br; br.addr = hdr + ldu64(hdr); br.top = 0; br.leaf = 0; br.cost = 64;
        fifo_push(&st, br);
      } else {                                                           /* BLAS leaf, :1076-1114 */
        LPUSH((uint64_t)(uintptr_t)next.addr + (uint64_t)off, 64);
        remaining -= 64; total_bvh += 64;                                /* quad and procedural both 64 */
        if (remaining < 0) { err = -8; break; }
      }
      /* ---- front-of-queue fit check (:1117-1268) ---- */
      if (st.head < st.tail && remaining - st.v[st.head].cost >= 0) { next = st.v[st.head++]; continue; }
      for (uint64_t i = st.head; i < st.tail; i++) fifo_push(&pend, st.v[i]);   /* :1167-1171 */
      break;
    }
    if (err) break;
    /* de-duplicate the node list, first occurrence wins (:1312-1330) */
    uint64_t w = list_first;
    for (uint64_t i = list_first; i < nln; i++) {
      int found = 0;
      for (uint64_t j = list_first; j < w; j++) if (ln[j].addr == ln[i].addr) { found = 1; break; }
      if (!found) ln[w++] = ln[i];
    }
    nln = w;
    if (ntl == captl) { captl = captl ? captl * 2 : 1024; tl = (vo_treelet*)realloc(tl, captl * sizeof(vo_treelet)); }
    tl[ntl].root = root_dev; tl[ntl].first = list_first; tl[ntl].count = (uint32_t)(nln - list_first); ntl++;
  }
#undef LPUSH
  free(pend.v); free(st.v); map_free(&done);
  if (err) { free(tl); free(ln); return err; }
  qsort(tl, ntl, sizeof(vo_treelet), cmp_treelet);     /* std::map iteration order = ascending root address */
  /* buildNodeToRootMap (:475-489): ascending roots, later roots overwrite -> highest root wins */
  map_init(&c->node_root, nln); map_init(&c->root_idx, ntl);
  for (uint64_t t = 0; t < ntl; t++) {
    map_put(&c->root_idx, tl[t].root, t);              /* treelet_addr_to_metadata_idx, :1332 */
    map_put(&c->node_root, tl[t].root, tl[t].root);
    for (uint32_t k = 0; k < tl[t].count; k++) map_put(&c->node_root, ln[tl[t].first + k].addr, tl[t].root);
  }
  c->treelets = tl; c->n_treelets = ntl; c->lnodes = ln; c->n_lnodes = nln; c->total_bvh_size = total_bvh;
  c->budget = budget; c->formed = 1;
  return 0;
}
uint64_t vo_treelet_count(vo_ctx* c) { return c->n_treelets; }
uint64_t vo_treelet_total_nodes(vo_ctx* c) { return c->n_lnodes; }
uint64_t vo_total_bvh_size(vo_ctx* c) { return c->total_bvh_size; }
void vo_treelet_table(vo_ctx* c, uint64_t* roots, uint32_t* counts, uint32_t* meta_idx, uint64_t* node_addr, uint32_t* node_size) {
  uint64_t k = 0;
  for (uint64_t t = 0; t < c->n_treelets; t++) {
    roots[t] = c->treelets[t].root; counts[t] = c->treelets[t].count; meta_idx[t] = (uint32_t)t;
    for (uint32_t i = 0; i < c->treelets[t].count; i++, k++) {
      node_addr[k] = c->lnodes[c->treelets[t].first + i].addr; node_size[k] = c->lnodes[c->treelets[t].first + i].size;
    }
  }
}
uint64_t vo_node_map_size(vo_ctx* c) { return c->node_root.n; }
static int cmp_u64pair(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
void vo_node_map(vo_ctx* c, uint64_t* nodes, uint64_t* roots) {
  uint64_t n = 0; uint64_t* tmp = (uint64_t*)malloc(c->node_root.n * 16);
  for (uint64_t i = 0; i < c->node_root.cap; i++) if (c->node_root.k[i] != ~0ull) { tmp[2 * n] = c->node_root.k[i]; tmp[2 * n + 1] = c->node_root.v[i]; n++; }
  qsort(tmp, n, 16, cmp_u64pair);
  for (uint64_t i = 0; i < n; i++) { nodes[i] = tmp[2 * i]; roots[i] = tmp[2 * i + 1]; }
  free(tmp);
}
int vo_addr_to_treelet(vo_ctx* c, uint64_t addr, uint64_t* root) { return map_get(&c->node_root, addr, root) ? 0 : -1; }
int vo_is_treelet_root(vo_ctx* c, uint64_t addr) { uint64_t v; return map_get(&c->root_idx, addr, &v); }

/* remapBVHToTreeletLayout, :1473-1509: treelet i (ascending root order) at base + i*(max+stride); the root
 * first, then the list entries that are not the root and not mapped yet, each advancing by its size even
 * when it was already mapped (:1501-1503). */
uint64_t vo_treelet_remap(vo_ctx* c, uint64_t base, uint32_t stride, uint64_t* orig, uint64_t* mapped) {
  map64 m; map_init(&m, c->n_lnodes);
  uint64_t n = 0;
  for (uint64_t t = 0; t < c->n_treelets; t++) {
    uint64_t root_new = base + t * ((uint64_t)c->budget + stride);
    uint64_t dummy;
    if (!map_get(&m, c->treelets[t].root, &dummy)) { map_put(&m, c->treelets[t].root, root_new); if (orig) { orig[n] = c->treelets[t].root; mapped[n] = root_new; } n++; }
    /* root.first.size: 64 for every root kind the reference creates (TLAS header 64 :847, internal 64, BLAS leaf 64)
       except an instance-leaf root, whose StackEntry carries size 128 (:927). */
    uint32_t root_size = 64;
    for (uint32_t i = 0; i < c->treelets[t].count; i++) if (c->lnodes[c->treelets[t].first + i].addr == c->treelets[t].root) root_size = c->lnodes[c->treelets[t].first + i].size;
    uint64_t cur = root_new + root_size;
    for (uint32_t i = 0; i < c->treelets[t].count; i++) {
      vo_lnode* e = &c->lnodes[c->treelets[t].first + i];
      if (e->addr == c->treelets[t].root) continue;
      if (!map_get(&m, e->addr, &dummy)) { map_put(&m, e->addr, cur); if (orig) { orig[n] = e->addr; mapped[n] = cur; } n++; }
      cur += e->size;
    }
  }
  map_free(&m);
  return n;
}

/* ---------------------------------------------------------------- traversal */
typedef struct { vo_txn* v; uint64_t n, cap; } txnbuf;
static void emit(txnbuf* b, uint64_t addr, uint32_t size, uint32_t type, uint64_t hist[9]) {
  if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 256; b->v = (vo_txn*)realloc(b->v, b->cap * sizeof(vo_txn)); }
  b->v[b->n].address = addr; b->v[b->n].size = size; b->v[b->n].type = type; b->n++;
  hist[type]++;
}
/* per-ray tree_level_map (std::map<uint8_t*, unsigned>, :1662): last write per address wins, max at the end */
typedef struct { uint64_t* k; uint32_t* v; uint32_t cap, n; } lvlmap;
static void lvl_put(lvlmap* m, uint64_t key, uint32_t val) {
  if ((m->n + 1) * 2 > m->cap) {
    uint32_t nc = m->cap ? m->cap * 2 : 64; uint64_t* nk = (uint64_t*)malloc(nc * 8); uint32_t* nv = (uint32_t*)malloc(nc * 4);
    memset(nk, 0xff, nc * 8);
    for (uint32_t i = 0; i < m->cap; i++) if (m->k[i] != ~0ull) { uint32_t j = (uint32_t)mix64(m->k[i]) & (nc - 1); while (nk[j] != ~0ull) j = (j + 1) & (nc - 1); nk[j] = m->k[i]; nv[j] = m->v[i]; }
    free(m->k); free(m->v); m->k = nk; m->v = nv; m->cap = nc;
  }
  uint32_t i = (uint32_t)mix64(key) & (m->cap - 1);
  while (m->k[i] != ~0ull && m->k[i] != key) i = (i + 1) & (m->cap - 1);
  if (m->k[i] == ~0ull) { m->k[i] = key; m->n++; }
  m->v[i] = val;
}
static uint32_t lvl_get(const lvlmap* m, uint64_t key) {
  uint32_t i = (uint32_t)mix64(key) & (m->cap - 1);
  while (m->k[i] != key) i = (i + 1) & (m->cap - 1);
  return m->v[i];
}
static void lvl_clear(lvlmap* m) { if (m->cap) memset(m->k, 0xff, (size_t)m->cap * 8); m->n = 0; }
static uint32_t lvl_max(const lvlmap* m) { uint32_t r = 0; for (uint32_t i = 0; i < m->cap; i++) if (m->k[i] != ~0ull && m->v[i] > r) r = m->v[i]; return r; }

/* StackEntry (vulkan_ray_tracing.h:180-208) reduced to what traversal reads: node, kind, instance context */
typedef struct { const uint8_t* addr; uint8_t top, leaf; int ictx; } sent;
typedef struct { rayf oray; float tmult; const uint8_t* leaf_addr; uint32_t instance_id; } ictx_t;
typedef struct { sent* v; uint32_t n, cap; } sstack;
static void spush(sstack* s, sent e) { if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->v = (sent*)realloc(s->v, s->cap * sizeof(sent)); } s->v[s->n++] = e; }

typedef struct {   /* per-thread scratch + accumulators */
  txnbuf tb; lvlmap lv; sstack cur, oth; ictx_t* ic; uint32_t nic, capic;
  vo_counters c;
} worker;

typedef struct {
  float min_thit, min_thit_object; int have; const uint8_t* leaf; ictx_t ictx;
} closest_t;

static void finish_ray(const vo_ray* r, const rayf* ray, const closest_t* cl, vo_hit* h, worker* w, uint32_t n_all_hits) {
  /* :2211-2245 / :2990-3033 */
  if (h) memset(h, 0, sizeof(*h));
  if (h) h->n_all_hits = n_all_hits;
  if (cl->min_thit < ray->tmax) {
    w->c.num_hits++;
    if (h) {
      h->hit = 1; h->t = cl->min_thit;
      h->geom = ldu32(cl->leaf + 4) & 0x0fffffffu; h->prim = ldu32(cl->leaf + 8); h->instance_id = cl->ictx.instance_id;
      for (int a = 0; a < 3; a++) h->point[a] = ray->o[a] + ray->d[a] * cl->min_thit;
      float p[3][3], op[3];
      for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) p[i][a] = ldf(cl->leaf + 16 + 12 * i + 4 * a);
      for (int a = 0; a < 3; a++) op[a] = cl->ictx.oray.o[a] + cl->ictx.oray.d[a] * cl->min_thit_object;
      barycentric(op, p[0], p[1], p[2], h->bary);
    }
  }
  (void)r;
}

/* quad leaf test shared by both variants (:2086-2165, :2802-2945); returns 1 if accepted */
static int quad_leaf(const uint8_t* leaf, const ictx_t* ic, float Tmin, float Tmax, float* t_obj, float* t_world, int* hit_out) {
  float p[3][3];
  for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) p[i][a] = ldf(leaf + 16 + 12 * i + 4 * a);
  float thit = 0.0f;
  int hit = ray_tri(p[0], p[1], p[2], &ic->oray, &thit);
  *hit_out = hit;
  if (!hit) return 0;
  float wt = thit / ic->tmult;
  *t_obj = thit; *t_world = wt;
  return (Tmin <= wt && wt <= Tmax);
}

/* traceRayWithTreelets, vulkan_ray_tracing.cc:1522-2307 */
static int trace_treelet(vo_ctx* c, const uint8_t* tlas, const vo_ray* r, vo_hit* h, worker* w) {
  const int64_t off = (int64_t)(c->tlas_dev - (uint64_t)(uintptr_t)tlas);
  const int terminate = (r->flags & 0x4u) != 0;
  if (terminate) w->c.n_anyhit_rays++; else w->c.n_closesthit_rays++;
  rayf ray; memcpy(ray.o, r->origin, 12); memcpy(ray.d, r->dir, 12); ray.tmin = r->tmin; ray.tmax = r->tmax;
  closest_t cl; memset(&cl, 0, sizeof(cl)); cl.min_thit = ray.tmax;
  uint32_t total_nodes = 0;
  txnbuf* tb = &w->tb; lvl_clear(&w->lv); w->cur.n = w->oth.n = 0; w->nic = 0;
  const uint64_t first_txn = tb->n;
  uint64_t* hist = w->c.mem_access_type;
#define DEV(p_) ((uint64_t)(uintptr_t)(p_) + (uint64_t)off)
  emit(tb, DEV(tlas), 64, T_STRUCT, hist);                                   /* :1685 */
  const uint8_t* top_root = tlas + ldu64(tlas);
  uint64_t cur_root = DEV(tlas);                                             /* :1707 */
  lvl_put(&w->lv, (uint64_t)(uintptr_t)top_root, 1);
  float idir[3]; calc_idir(ray.d, idir);
  {
    float lo[3] = { ldf(tlas + 8), ldf(tlas + 12), ldf(tlas + 16) }, hi[3] = { ldf(tlas + 20), ldf(tlas + 24), ldf(tlas + 28) }, th;
    if (ray_box(lo, hi, idir, ray.o, ray.tmin, ray.tmax, &th)) {             /* :1722 */
      uint64_t ct; if (!map_get(&c->node_root, DEV(top_root), &ct)) return -6;
      sent e = { top_root, 1, 0, -1 };
      spush(cur_root == ct ? &w->cur : &w->oth, e);
    }
  }
  while (w->cur.n || w->oth.n) {
    if (!w->cur.n) {                                                         /* :1748-1754: host address stored */
      sent e = w->oth.v[--w->oth.n]; spush(&w->cur, e); cur_root = (uint64_t)(uintptr_t)e.addr;
    }
    sent n = w->cur.v[--w->cur.n];
    if (!n.leaf) {                                                           /* internal, TLAS :1759 / BLAS :1954 */
      inode nd; unpack_internal(&nd, n.addr);
      emit(tb, DEV(n.addr), 64, T_INTERNAL, hist); total_nodes++;
      const rayf* rr = n.top ? &ray : &w->ic[n.ictx].oray;
      const float tm = n.top ? 1.0f : w->ic[n.ictx].tmult;
      float id[3]; calc_idir(rr->d, id);
      int hit[6];
      for (int i = 0; i < 6; i++) {
        hit[i] = 0;
        if (nd.size[i] > 0) {
          float lo[3], hi[3], th; child_bounds(&nd, i, lo, hi);
          hit[i] = ray_box(lo, hi, id, rr->o, rr->tmin, rr->tmax, &th);
          if (n.top) { if (hit[i] && th >= cl.min_thit) hit[i] = 0; }          /* :1791 */
          else { if (hit[i] && th >= cl.min_thit * tm) hit[i] = 0; }           /* :1989 */
        }
      }
      const uint8_t* child = n.addr + (int64_t)nd.child_offset * 64;
      uint32_t plevel = lvl_get(&w->lv, (uint64_t)(uintptr_t)n.addr);
      for (int i = 0; i < 6; i++) {
        if (hit[i]) {
          sent e; e.addr = child; e.top = n.top; e.ictx = n.ictx;
          if (nd.type[i] != NODE_INTERNAL) { if (n.top && nd.type[i] != NODE_INSTANCE) return -6; e.leaf = 1; } else e.leaf = 0;
          uint64_t ct; if (!map_get(&c->node_root, DEV(child), &ct)) return -6;   /* addrToTreeletID assert :470 */
          spush(cur_root == ct ? &w->cur : &w->oth, e);                        /* push_front: later slots pop first */
          lvl_put(&w->lv, (uint64_t)(uintptr_t)child, plevel + 1);
        }
        child += nd.size[i] * 64;
      }
    } else if (n.top) {                                                      /* instance leaf, :1876-1953 */
      ileaf l; unpack_instance(&l, n.addr);
      emit(tb, DEV(n.addr), 128, T_INSTANCE, hist); total_nodes++;
      const uint8_t* hdr = n.addr + l.bvh_address;
      uint64_t hdr_dev; if (!map_get(&c->blas, (uint64_t)(uintptr_t)hdr, &hdr_dev)) return -5;
      emit(tb, hdr_dev, 64, T_STRUCT, hist);                                  /* :1913 */
      if (w->nic == w->capic) { w->capic = w->capic ? w->capic * 2 : 16; w->ic = (ictx_t*)realloc(w->ic, w->capic * sizeof(ictx_t)); }
      ictx_t* ic = &w->ic[w->nic];
      transform_ray(&ray, l.w2o, &ic->oray, &ic->tmult); ic->leaf_addr = n.addr; ic->instance_id = l.instance_id;
      const uint8_t* broot = hdr + ldu64(hdr);
      uint64_t ct; if (!map_get(&c->node_root, DEV(broot), &ct)) return -6;
      sent e = { broot, 0, 0, (int)w->nic }; w->nic++;
      spush(cur_root == ct ? &w->cur : &w->oth, e);
      lvl_put(&w->lv, (uint64_t)(uintptr_t)broot, lvl_get(&w->lv, (uint64_t)(uintptr_t)n.addr));   /* :1944 */
    } else {                                                                 /* BLAS leaf, :2073-2204 */
      emit(tb, DEV(n.addr), 8, T_DESC, hist);
      uint32_t dw = ldu32(n.addr + 4);
      if (((dw >> 29) & 1) == 0) {
        if ((ldu32(n.addr + 12) & 0x1ffff) != 0) return -6;                   /* assert PrimitiveIndex1Delta == 0, :2108 */
        const ictx_t* ic = &w->ic[n.ictx];
        float to, tw; int hit;
        int acc = quad_leaf(n.addr, ic, r->tmin, r->tmax, &to, &tw, &hit) && tw < cl.min_thit;   /* :2127 */
        if (acc) {
          cl.min_thit = tw; cl.min_thit_object = to; cl.leaf = n.addr; cl.ictx = *ic; cl.have = 1;
          emit(tb, DEV(n.addr), 64, T_QUAD_HIT, hist); total_nodes++;
          if (terminate) { w->cur.n = 0; w->oth.n = 0; }
        } else { emit(tb, DEV(n.addr), 64, T_QUAD, hist); total_nodes++; }
      } else {
        emit(tb, DEV(n.addr), 64, T_PROC, hist); total_nodes++;               /* intersection-table call: vo_table_events */
        proc_visit(c, w->ic[n.ictx].leaf_addr);
      }
    }
  }
#undef DEV
  finish_ray(r, &ray, &cl, h, w, 0);
  for (uint64_t i = first_txn; i < tb->n; i++) w->c.accessed_data_size += tb->v[i].size;   /* :2257-2261 */
  if (total_nodes > w->c.max_nodes_per_ray) w->c.max_nodes_per_ray = total_nodes;
  w->c.tot_nodes_per_ray += total_nodes;
  uint32_t lv = lvl_max(&w->lv); if (lv > w->c.max_tree_depth) w->c.max_tree_depth = lv;
  return 0;
}

/* traceRay, vulkan_ray_tracing.cc:2309-3076 */
static int trace_dfs(vo_ctx* c, const uint8_t* tlas, const vo_ray* r, vo_hit* h, worker* w) {
  const int64_t tlas_off = (int64_t)(c->tlas_dev - (uint64_t)(uintptr_t)tlas);
  int64_t off = tlas_off;
  const int terminate = (r->flags & 0x4u) != 0, opaque = (r->flags & 0x1u) != 0;
  if (terminate) w->c.n_anyhit_rays++; else w->c.n_closesthit_rays++;
  rayf ray; memcpy(ray.o, r->origin, 12); memcpy(ray.d, r->dir, 12); ray.tmin = r->tmin; ray.tmax = r->tmax;
  closest_t cl; memset(&cl, 0, sizeof(cl)); cl.min_thit = ray.tmax;
  uint32_t total_nodes = 0, n_all_hits = 0;
  txnbuf* tb = &w->tb; lvl_clear(&w->lv); sstack* st = &w->cur; st->n = 0;
  uint64_t* hist = w->c.mem_access_type;
#define DEV(p_) ((uint64_t)(uintptr_t)(p_) + (uint64_t)off)
  emit(tb, DEV(tlas), 64, T_STRUCT, hist);
  const uint8_t* top_root = tlas + ldu64(tlas);
  lvl_put(&w->lv, (uint64_t)(uintptr_t)top_root, 1);
  float idir[3]; calc_idir(ray.d, idir);
  {
    float lo[3] = { ldf(tlas + 8), ldf(tlas + 12), ldf(tlas + 16) }, hi[3] = { ldf(tlas + 20), ldf(tlas + 24), ldf(tlas + 28) }, th;
    if (ray_box(lo, hi, idir, ray.o, ray.tmin, ray.tmax, &th)) { sent e = { top_root, 1, 0, -1 }; spush(st, e); }
  }
  while (st->n) {
    const uint8_t* next = NULL;
    if (!st->v[st->n - 1].leaf) next = st->v[--st->n].addr;                   /* :2494-2498 */
    while (next) {                                                           /* TLAS internal chain, :2500-2599 */
      off = tlas_off;
      const uint8_t* na = next; next = NULL;
      inode nd; unpack_internal(&nd, na);
      emit(tb, DEV(na), 64, T_INTERNAL, hist); total_nodes++;
      int hit[6];
      for (int i = 0; i < 6; i++) {
        hit[i] = 0;
        if (nd.size[i] > 0) { float lo[3], hi[3], th; child_bounds(&nd, i, lo, hi);
          hit[i] = ray_box(lo, hi, idir, ray.o, ray.tmin, ray.tmax, &th); if (hit[i] && th >= cl.min_thit) hit[i] = 0; }
      }
      const uint8_t* child = na + (int64_t)nd.child_offset * 64;
      uint32_t plevel = lvl_get(&w->lv, (uint64_t)(uintptr_t)na);
      for (int i = 0; i < 6; i++) {
        if (hit[i]) {
          if (nd.type[i] != NODE_INTERNAL) { if (nd.type[i] != NODE_INSTANCE) return -6; sent e = { child, 1, 1, -1 }; spush(st, e); }
          else if (!next) next = child;                                      /* first hit internal child is followed, :2573 */
          else { sent e = { child, 1, 0, -1 }; spush(st, e); }
          lvl_put(&w->lv, (uint64_t)(uintptr_t)child, plevel + 1);
        }
        child += nd.size[i] * 64;
      }
    }
    while (st->n && st->v[st->n - 1].leaf) {                                  /* TLAS leaves, :2602-2987 */
      off = tlas_off;
      const uint8_t* la = st->v[--st->n].addr;
      ileaf l; unpack_instance(&l, la);
      emit(tb, DEV(la), 128, T_INSTANCE, hist); total_nodes++;
      const uint8_t* hdr = la + l.bvh_address;
      uint64_t hdr_dev; if (!map_get(&c->blas, (uint64_t)(uintptr_t)hdr, &hdr_dev)) return -5;
      off = (int64_t)(hdr_dev - (uint64_t)(uintptr_t)hdr);                    /* :2640: BLAS offset from here on */
      emit(tb, DEV(hdr), 64, T_STRUCT, hist);
      ictx_t ic; transform_ray(&ray, l.w2o, &ic.oray, &ic.tmult); ic.leaf_addr = la; ic.instance_id = l.instance_id;
      float oid[3]; calc_idir(ic.oray.d, oid);
      const uint8_t* broot = hdr + ldu64(hdr);
      { sent e = { broot, 0, 0, 0 }; spush(st, e); }
      lvl_put(&w->lv, (uint64_t)(uintptr_t)broot, lvl_get(&w->lv, (uint64_t)(uintptr_t)la));
      while (st->n && !st->v[st->n - 1].top) {                                /* :2679 */
        const uint8_t* nx = st->v[--st->n].addr;
        while (nx) {                                                         /* BLAS internal chain, :2687-2786 */
          const uint8_t* na = nx; nx = NULL;
          inode nd; unpack_internal(&nd, na);
          emit(tb, DEV(na), 64, T_INTERNAL, hist); total_nodes++;
          int hit[6];
          for (int i = 0; i < 6; i++) {
            hit[i] = 0;
            if (nd.size[i] > 0) { float lo[3], hi[3], th; child_bounds(&nd, i, lo, hi);
              hit[i] = ray_box(lo, hi, oid, ic.oray.o, ic.oray.tmin, ic.oray.tmax, &th);
              if (hit[i] && th >= cl.min_thit * ic.tmult) hit[i] = 0; }        /* :2725 */
          }
          const uint8_t* child = na + (int64_t)nd.child_offset * 64;
          uint32_t plevel = lvl_get(&w->lv, (uint64_t)(uintptr_t)na);
          for (int i = 0; i < 6; i++) {
            if (hit[i]) {
              if (nd.type[i] != NODE_INTERNAL) { sent e = { child, 0, 1, 0 }; spush(st, e); }
              else if (!nx) nx = child;
              else { sent e = { child, 0, 0, 0 }; spush(st, e); }
              lvl_put(&w->lv, (uint64_t)(uintptr_t)child, plevel + 1);
            }
            child += nd.size[i] * 64;
          }
        }
        while (st->n && !st->v[st->n - 1].top && st->v[st->n - 1].leaf) {      /* BLAS leaves, :2789-2985 */
          const uint8_t* lf = st->v[--st->n].addr;
          emit(tb, DEV(lf), 8, T_DESC, hist);
          uint32_t dw = ldu32(lf + 4);
          if (((dw >> 29) & 1) == 0) {
            if ((ldu32(lf + 12) & 0x1ffff) != 0) return -6;
            float to, tw; int hit;
            if (quad_leaf(lf, &ic, r->tmin, r->tmax, &to, &tw, &hit)) {        /* no `< min_thit` test, :2843 */
              if (opaque && tw < cl.min_thit) cl.min_thit = tw;               /* :2850-2852 */
              cl.min_thit_object = to; cl.leaf = lf; cl.ictx = ic; cl.have = 1; /* overwritten by EVERY accepted hit */
              emit(tb, DEV(lf), 64, T_QUAD_HIT, hist); total_nodes++;
              if (!opaque) { w->c.num_any_hits++; n_all_hits++; }              /* any-hit table txns: not restated */
              if (terminate) st->n = 0;
            } else { emit(tb, DEV(lf), 64, T_QUAD, hist); total_nodes++; }
          } else { emit(tb, DEV(lf), 64, T_PROC, hist); total_nodes++; proc_visit(c, ic.leaf_addr); }
        }
      }
    }
  }
#undef DEV
  finish_ray(r, &ray, &cl, h, w, n_all_hits);
  if (total_nodes > w->c.max_nodes_per_ray) w->c.max_nodes_per_ray = total_nodes;
  w->c.tot_nodes_per_ray += total_nodes;
  uint32_t lv = lvl_max(&w->lv); if (lv > w->c.max_tree_depth) w->c.max_tree_depth = lv;
  return 0;
}

static void worker_free(worker* w) { free(w->tb.v); free(w->lv.k); free(w->lv.v); free(w->cur.v); free(w->oth.v); free(w->ic); }
static void merge_counters(vo_counters* d, const vo_counters* s) {
  for (int i = 0; i < 9; i++) d->mem_access_type[i] += s->mem_access_type[i];
  d->num_hits += s->num_hits; d->num_any_hits += s->num_any_hits; d->n_anyhit_rays += s->n_anyhit_rays;
  d->n_closesthit_rays += s->n_closesthit_rays; d->tot_nodes_per_ray += s->tot_nodes_per_ray;
  d->accessed_data_size += s->accessed_data_size;
  if (s->max_nodes_per_ray > d->max_nodes_per_ray) d->max_nodes_per_ray = s->max_nodes_per_ray;
  if (s->max_tree_depth > d->max_tree_depth) d->max_tree_depth = s->max_tree_depth;
}

/* Batch driver.  mode 0 = traceRay, 1 = traceRayWithTreelets.  Returns #transactions (negated if `cap` was
 * too small, or a small negative error code -5/-6/-8 shifted below -2^40 never occurs: errors return
 * INT64_MIN + code).  nthreads > 1 shards contiguous ray blocks over OpenMP threads (rays are independent;
 * the reference itself is strictly sequential, abstract_hardware_model.cc:3052-3063). */
int64_t vo_trace(vo_ctx* c, const void* tlas_v, int mode, uint32_t n, const vo_ray* rays, vo_hit* hits, uint32_t* counts,
                 vo_txn* txns, uint64_t cap, uint64_t* treelet_ids, int nthreads) {
  const uint8_t* tlas = (const uint8_t*)tlas_v;
  if (!c->formed) return INT64_MIN + 1;
  if (nthreads < 1) nthreads = 1;
  uint32_t* cnt = counts ? counts : (uint32_t*)malloc((size_t)n * 4 + 4);
  worker* ws = (worker*)calloc((size_t)nthreads, sizeof(worker));
  uint64_t* blk_first = (uint64_t*)calloc((size_t)nthreads + 1, 8);
  int err = 0;
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int t = 0, nt = 1;
#endif
    if (t < nthreads) {
      worker* w = &ws[t];
      uint32_t lo = (uint32_t)((uint64_t)n * t / nt), hi = (uint32_t)((uint64_t)n * (t + 1) / nt);
      for (uint32_t i = lo; i < hi; i++) {
        uint64_t before = w->tb.n;
        int e = mode == 1 ? trace_treelet(c, tlas, &rays[i], hits ? &hits[i] : NULL, w) : trace_dfs(c, tlas, &rays[i], hits ? &hits[i] : NULL, w);
        if (e) {
#pragma omp atomic write
          err = e;
          break;
        }
        cnt[i] = (uint32_t)(w->tb.n - before);
      }
      blk_first[t + 1] = w->tb.n;
    }
  }
  int64_t ret;
  if (err) ret = INT64_MIN - (int64_t)err;
  else {
    for (int t = 0; t < nthreads; t++) blk_first[t + 1] += blk_first[t];
    uint64_t total = blk_first[nthreads];
    if (txns) {
      for (int t = 0; t < nthreads; t++) {
        uint64_t first = blk_first[t], m = ws[t].tb.n;
        if (first >= cap) break;
        if (first + m > cap) m = cap - first;
        memcpy(txns + first, ws[t].tb.v, m * sizeof(vo_txn));
      }
      if (treelet_ids) {
        uint64_t m = total < cap ? total : cap;
        for (uint64_t i = 0; i < m; i++) { uint64_t rt; treelet_ids[i] = map_get(&c->node_root, txns[i].address, &rt) ? rt : ~0ull; }
      }
    }
    ret = (txns && total > cap) ? -(int64_t)total : (int64_t)total;
    for (int t = 0; t < nthreads; t++) merge_counters(&c->c, &ws[t].c);
    c->c.ray_count += n;
  }
  for (int t = 0; t < nthreads; t++) worker_free(&ws[t]);
  free(ws); free(blk_first); if (!counts) free(cnt);
  return ret;
}

/* ================================================================ RT-unit replay helpers
 * Restatement of the two pure pieces of rt_unit (gpgpu-sim/shader.cc) that consume the trace; pinned against the
 * reference bodies in oracle/_ref (tests/test_oracle.py). */

/* rt_unit::sort_mem_accesses, shader.cc:3012-3089, on one ray's list t[0..n), in place.
 * method 1 ("loose", :3057-3086): treelets in order of first appearance; inside a treelet the accesses in their original
 *   order -- but every access is replaced by the FIRST record of the list that has the same address (:3069-3076), so the
 *   8-byte descriptor record of a leaf appears twice and its 64-byte record never.
 * method 0 ("strict", :3027-3056): treelets in order of first appearance; for each, walk the treelet's node list in
 *   formation order and emit all records of that address, original order (:3046-3054).  A node listed by several treelets
 *   is emitted with the first visited treelet that lists it. */
static void sort_one(vo_ctx* c, int method, vo_txn* t, uint32_t n, vo_txn* tmp, uint64_t* tag, uint64_t* order, uint8_t* done) {
  uint32_t n_order = 0, o = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (!map_get(&c->node_root, t[i].address, &tag[i])) tag[i] = 0;   /* addrToTreeletID; the reference asserts on unknown addresses */
    uint32_t j = 0; while (j < n_order && order[j] != tag[i]) j++;
    if (j == n_order) order[n_order++] = tag[i];
  }
  if (method == 1) {
    for (uint32_t q = 0; q < n_order; q++)
      for (uint32_t i = 0; i < n; i++) if (tag[i] == order[q]) {
        uint32_t f = 0; while (t[f].address != t[i].address) f++;
        tmp[o++] = t[f];
      }
  } else {
    memset(done, 0, n);
    for (uint32_t q = 0; q < n_order; q++) {
      uint64_t idx;
      if (!map_get(&c->root_idx, order[q], &idx)) continue;
      const vo_treelet* tl = &c->treelets[idx];
      for (uint32_t k = 0; k < tl->count; k++) {
        const uint64_t a = c->lnodes[tl->first + k].addr;
        for (uint32_t i = 0; i < n; i++) if (!done[i] && t[i].address == a) { tmp[o++] = t[i]; done[i] = 1; }
      }
    }
  }
  /* the reference asserts mem_accesses.size() == sorted.size() (:3087); keep whatever was not placed at the end */
  if (method == 0) for (uint32_t i = 0; i < n && o < n; i++) if (!done[i]) tmp[o++] = t[i];
  memcpy(t, tmp, (size_t)n * sizeof(vo_txn));
}
void vo_sort_trace(vo_ctx* c, int method, uint64_t n_rays, const uint64_t* offsets, vo_txn* txns) {
  uint32_t cap = 0;
  for (uint64_t r = 0; r < n_rays; r++) { const uint64_t n = offsets[r + 1] - offsets[r]; if (n > cap) cap = (uint32_t)n; }
  vo_txn* tmp = (vo_txn*)malloc((size_t)(cap + 1) * sizeof(vo_txn));
  uint64_t* tag = (uint64_t*)malloc((size_t)(cap + 1) * 8); uint64_t* order = (uint64_t*)malloc((size_t)(cap + 1) * 8);
  uint8_t* done = (uint8_t*)malloc(cap + 1);
  for (uint64_t r = 0; r < n_rays; r++) sort_one(c, method, txns + offsets[r], (uint32_t)(offsets[r + 1] - offsets[r]), tmp, tag, order, done);
  free(tmp); free(tag); free(order); free(done);
}

/* The treelet-prefetch vote of rt_unit::cycle, shader.cc:3419-3640, for one group of rays with a fresh unit (no
 * last_prefetched_treelet history).  Ray r votes with addrToTreeletID of its pending access txns[offsets[r] + front[r]]
 * (:3424-3433); the winner is the first maximum in ascending root-address order (std::map iteration, strict '>').
 *   heuristic 0: always submit the whole treelet (:3438-3446)
 *   heuristic 1: submit iff votes/total >= threshold (:3487-3514)
 *   heuristic 2: submit the first (int)(n_nodes * votes/total + 0.5) nodes (:3515-3534)
 *   heuristic 3: submit the last that many nodes (:3535-3552, :3566)
 * Chunks (:3566-3620): per node, when load_metadata, per_meta/32 chunks of the metadata row
 * treelet_addr_to_metadata_idx[NODE address] (0 unless the node is itself a root -- std::map::operator[], :3571) first,
 * then ceil(size/32) chunks of the node; each chunk is (address, owner address). */
typedef struct { uint64_t root; uint32_t votes, total, submit, n_nodes, first_node, num_nodes; } vo_prefetch_decision;
int64_t vo_prefetch_vote(vo_ctx* c, int heuristic, double threshold, int load_metadata, uint64_t metadata_base, uint32_t per_meta,
                         uint64_t n_rays, const uint64_t* ray_ids, const uint64_t* offsets, const uint32_t* front, const vo_txn* txns,
                         vo_prefetch_decision* dec, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t cap) {
  uint32_t* tally = (uint32_t*)calloc(c->n_treelets ? c->n_treelets : 1, 4);
  uint32_t total = 0;
  for (uint64_t i = 0; i < n_rays; i++) {
    const uint64_t r = ray_ids ? ray_ids[i] : i, k = offsets[r] + (front ? front[r] : 0);
    uint64_t root, idx;
    if (k >= offsets[r + 1]) continue;
    if (!map_get(&c->node_root, txns[k].address, &root) || !map_get(&c->root_idx, root, &idx)) continue;
    tally[idx]++; total++;
  }
  int64_t best = -1; uint32_t bestv = 0;
  for (uint64_t t = 0; t < c->n_treelets; t++) if (tally[t] > bestv) { bestv = tally[t]; best = (int64_t)t; }   /* 0-vote treelets are not in the map */
  memset(dec, 0, sizeof(*dec));
  dec->total = total;
  uint64_t n = 0;
  if (best >= 0) {
    const vo_treelet* tl = &c->treelets[best];
    const double pct = (double)bestv / (double)total;
    dec->root = tl->root; dec->votes = bestv; dec->n_nodes = tl->count;
    const uint32_t part = (uint32_t)(int)((double)tl->count * pct + 0.5);
    dec->submit = (heuristic == 1) ? (pct >= threshold) : 1u;
    dec->num_nodes = (heuristic == 2 || heuristic == 3) ? part : tl->count;
    dec->first_node = (heuristic == 3) ? tl->count - part : 0u;
    if (dec->submit) {
      for (uint32_t j = dec->first_node; j < dec->first_node + dec->num_nodes; j++) {
        const vo_lnode* e = &c->lnodes[tl->first + j];
        if (load_metadata) {
          uint64_t idx = 0; if (!map_get(&c->root_idx, e->addr, &idx)) idx = 0;
          const uint64_t ma = metadata_base + idx * (uint64_t)per_meta;
          for (uint32_t q = 0; q < per_meta / 32; q++) { if (n < cap && chunk_addr) { chunk_addr[n] = ma + q * 32ull; chunk_owner[n] = ma; } n++; }
        }
        for (uint32_t q = 0; q < (e->size + 31) / 32; q++) { if (n < cap && chunk_addr) { chunk_addr[n] = e->addr + q * 32ull; chunk_owner[n] = e->addr; } n++; }
      }
    }
  }
  free(tally);
  return (int64_t)n;
}

/* rt_unit::schedule_next_warp, shader.cc:4307-4392, for one RT unit.  Warps in m_current_warps order; lane l of warp w is
 * ray ray_ids[32 * w + l] (~0 = no thread); a thread "matches" when addrToTreeletID of its pending access equals
 * last_prefetched_treelet.  scheduler 1: first non-stalled warp with a matching thread (:4313-4343); scheduler 2: the
 * non-stalled warp with the most matching threads, first among equals, at least one (:4345-4378); both fall back to --
 * and scheduler 0 is -- the first non-stalled warp (:4381-4390).  Returns the warp index or -1. */
int64_t vo_schedule_pick(vo_ctx* c, int scheduler, uint64_t last_prefetched, uint64_t n_warps, const uint64_t* ray_ids, const uint8_t* stalled,
                         const uint64_t* offsets, const uint32_t* front, const vo_txn* txns) {
  int64_t first_free = -1, best = -1; uint32_t best_n = 0;
  for (uint64_t w = 0; w < n_warps; w++) {
    if (stalled && stalled[w]) continue;
    if (first_free < 0) first_free = (int64_t)w;
    if (scheduler != 1 && scheduler != 2) break;
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) {
      const uint64_t r = ray_ids[32 * w + l]; uint64_t root;
      if (r == ~0ull) continue;
      const uint64_t k = offsets[r] + (front ? front[r] : 0);
      if (k >= offsets[r + 1]) continue;
      if (map_get(&c->node_root, txns[k].address, &root) && root == last_prefetched) m++;
    }
    if (scheduler == 1 && m) return (int64_t)w;
    if (scheduler == 2 && m > best_n) { best_n = m; best = (int64_t)w; }
  }
  return best >= 0 ? best : first_free;
}

/* ================================================================ shader-table side effects, Baseline tables
 * Every procedural-leaf visit (both variants, :2171-2203 / :2951-2984) calls intersection_table[cta]->add_intersection, every
 * accepted triangle hit of a NON-opaque ray in traceRay (:2869-2930) calls anyhit_table[cta]->add_intersection and pushes a
 * Hit_data.  The Baseline table (intersection_table.cc:165-187) keeps a row counter per thread (index[tid]) and emits two
 * stores: &table[row].hitGroupIndex[tid] (4 bytes) and &table[row].shader_data[tid] (8 bytes); no loads, so the load trace is
 * unaffected.  Restated from the trace: rays [32g, 32g+32) are the threads of one CTA row (tid = tid_x[r] or r % 32), tables
 * empty at the start of the batch.  Uniform host->device offset only (the record addresses are mapped back with it). */
typedef struct { uint32_t table, shader_counter, hit_group_index, primitive_id, instance_id, tid; uint64_t store_addr[2]; uint32_t store_size[2]; } vo_table_event;
enum { VO_BASELINE_ENTRY = 384 };   /* sizeof(Baseline_Entry): 32 x u32 + 32 x {u32, u32}, intersection_table.h:109-116 */
int64_t vo_table_events(vo_ctx* c, const void* tlas_v, int mode, uint64_t n_rays, const vo_ray* rays, const uint64_t* offsets, const vo_txn* txns,
                        const uint8_t* tid_x, const uint64_t* proc_inst, uint64_t itab_base, uint64_t atab_base, uint32_t* ev_counts, vo_table_event* ev, vo_hit* anyhit, uint64_t cap) {
  const int64_t off = (int64_t)(c->tlas_dev - (uint64_t)(uintptr_t)tlas_v);
  uint32_t rows[2][32];
  uint64_t total = 0, n_proc = 0;   /* proc_inst: the instance leaf of every procedural visit, in trace order (vo_set_proc_sink) -- the
                                       treelet-ordered variant interleaves the nodes of different instances, so the trace alone does not tell */
  for (uint64_t r = 0; r < n_rays; r++) {
    if (r % 32 == 0) memset(rows, 0, sizeof(rows));
    const uint32_t tid = tid_x ? tid_x[r] : (uint32_t)(r % 32);
    const uint8_t* inst = NULL; uint32_t n_ev = 0;
    for (uint64_t k = offsets[r]; k < offsets[r + 1]; k++) {
      const uint8_t* node = (const uint8_t*)(uintptr_t)(txns[k].address - (uint64_t)off);
      if (txns[k].type == T_INSTANCE) { inst = node; continue; }
      int table = -1;
      if (txns[k].type == T_PROC) table = 0;
      else if (txns[k].type == T_QUAD_HIT && mode == 0 && !(rays[r].flags & 0x1u)) table = 1;
      if (table < 0) continue;
      const uint8_t* own = table == 0 ? (const uint8_t*)(uintptr_t)proc_inst[n_proc++] : inst;
      if (!own) continue;
      ileaf il; unpack_instance(&il, own);
      if (total < cap && ev) {
        vo_table_event* e = &ev[total];
        const uint64_t base = table ? atab_base : itab_base;
        e->table = (uint32_t)table; e->shader_counter = rows[table][tid]; e->tid = tid;
        e->hit_group_index = il.hit_group; e->instance_id = il.instance_id;
        e->primitive_id = table ? ldu32(node + 8) : ldu32(node + 12);                  /* PrimitiveIndex0 / PrimitiveIndex[0] */
        e->store_addr[0] = base + (uint64_t)e->shader_counter * VO_BASELINE_ENTRY + 4ull * tid; e->store_size[0] = 4;
        e->store_addr[1] = base + (uint64_t)e->shader_counter * VO_BASELINE_ENTRY + 128ull + 8ull * tid; e->store_size[1] = 8;
        if (anyhit) {
          vo_hit* h = &anyhit[total]; memset(h, 0, sizeof(*h));
          if (table == 1) {
            /* the any-hit Hit_data (:2886-2922): same transform, triangle test and barycentrics as the traversal itself */
            rayf wr, orr; float tmult, thit = 0.0f, p[3][3], op[3];
            for (int a = 0; a < 3; a++) { wr.o[a] = rays[r].origin[a]; wr.d[a] = rays[r].dir[a]; }
            wr.tmin = rays[r].tmin; wr.tmax = rays[r].tmax;
            transform_ray(&wr, il.w2o, &orr, &tmult);
            for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) p[i][a] = ldf(node + 16 + 12 * i + 4 * a);
            ray_tri(p[0], p[1], p[2], &orr, &thit);
            const float tw = thit / tmult;
            h->hit = 1; h->t = tw; h->geom = ldu32(node + 4) & 0x0fffffffu; h->prim = ldu32(node + 8); h->instance_id = il.instance_id;
            for (int a = 0; a < 3; a++) { h->point[a] = wr.o[a] + wr.d[a] * tw; op[a] = orr.o[a] + orr.d[a] * thit; }
            barycentric(op, p[0], p[1], p[2], h->bary);
          }
        }
      }
      rows[table][tid]++; n_ev++; total++;
    }
    if (ev_counts) ev_counts[r] = n_ev;
  }
  return (int64_t)total;
}

/* ---- Function_Call_Coalescing intersection table (-gpgpu_rt_intersection_table_type 1).
 * Coalescing_warp_intersection_table::add_intersection, intersection_table.cc:43-98: the rows are shared by the 32 threads of a
 * CTA; a call scans rows 0.. (one Intersection_Table_Load record per row looked at), claims the first row of its hit group
 * whose thread_mask[tid] is free (two stores) or appends a row (three stores).  The caller merges the loads into the ray's
 * transaction list unless the address is already there (vulkan_ray_tracing.cc:2186-2200 / :2966-2980), i.e. only the rows the
 * ray has not looked at before.  Input: the table-0 events of a batch in ray order (rays [32g, 32g + 32) = one CTA, fresh
 * table per CTA); output per event: row, appended, n_loads, first_new_load.  Returns -1 if a CTA needs more than the 100 rows
 * the reference allocates (INTERSECTION_TABLE_MAX_LENGTH; it would write past its allocation). */
typedef struct { uint32_t row, appended, n_loads, first_new_load; } vo_coalescing_event;
enum { VO_COALESCING_ROWS = 100 };
int vo_coalescing_events(uint64_t n_rays, const uint64_t* event_offsets, const vo_table_event* ev, vo_coalescing_event* out) {
  uint32_t key[VO_COALESCING_ROWS], mask[VO_COALESCING_ROWS], n_rows = 0;
  for (uint64_t r = 0; r < n_rays; r++) {
    if (r % 32 == 0) n_rows = 0;
    uint32_t seen = 0;                                   /* rows whose load record is already in this ray's list */
    for (uint64_t k = event_offsets[r]; k < event_offsets[r + 1]; k++) {
      vo_coalescing_event* o = &out[k];
      o->row = o->appended = o->n_loads = o->first_new_load = 0;
      if (ev[k].table != 0) continue;                    /* the any-hit table is a Baseline table whatever the option says (:441-447) */
      const uint32_t bit = 1u << (ev[k].tid & 31u);
      uint32_t i = 0, found = 0;
      for (; i < n_rows; i++) if (key[i] == ev[k].hit_group_index && !(mask[i] & bit)) { found = 1; break; }
      if (found) { mask[i] |= bit; o->row = i; o->n_loads = i + 1; }
      else {
        if (n_rows >= VO_COALESCING_ROWS) return -1;
        key[n_rows] = ev[k].hit_group_index; mask[n_rows] = bit; o->row = n_rows; o->appended = 1; o->n_loads = n_rows; n_rows++;
      }
      o->first_new_load = seen < o->n_loads ? seen : o->n_loads;
      if (o->n_loads > seen) seen = o->n_loads;
    }
  }
  return 0;
}
