// traverse_tb.cu -- K1, treelet-binned wavefront variant (north_star item 2; VSRT_K1_TB=1 / vsrt_config.k1_variant).
//
//   VulkanRayTracing::traceRayWithTreelets (vulkan_ray_tracing.cc:1522-2307): a ray drains the nodes of its CURRENT treelet
//   (current_treelet_stack, :1743-1756) and only then takes the front of other_treelet_stack and makes that node's treelet the
//   current one (:1748-1754, :1832-1856).
//
// A ray's own visit order is fixed by that rule, but between two switches it touches one treelet only -- so the batch can be run
// as ROUNDS: in every round each live ray is binned by the treelet it enters next, the bins are made contiguous by a radix sort
// of (treelet, ray) pairs, and a CTA of 128 consecutive rays of the sorted list
//   * stages the treelet most of its rays are in into shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier) from a
//     treelet-layout copy of the arena (every treelet's nodes contiguous, in the list order of createTreelets -- the layout the
//     reference's remapBVHToTreeletLayout describes, :1473-1509), if at least TB_STAGE_MIN_RAYS of them share it;
//   * groups its lanes by treelet with __match_any_sync (one lane per group fetches the treelet's layout record);
//   * lets every lane drain its ray's current list -- node bytes come from the staged copy for rays in the staged treelet, from
//     the arena otherwise -- pushing out-of-treelet children onto the ray's `other` stack in global memory;
//   * writes the ray's state back and files it under its next treelet.
// Per-ray results are bit-identical to the lane-owned kernel (traverse.cu) by construction: same entries, same pops, same tests.
// Rays that need the EXACT slab test are left to that kernel's EXACT pass (counts[r] = RAY_DEFERRED), as there.
//
// Inside a staged treelet a node is found by POSITION, not by address: K0's lists are BFS order, so the children of a node that
// stay in its treelet are a prefix of its present children and sit next to each other in the list.  The layout copy carries,
// in bytes the reference never reads, the position P of a node's first in-treelet child (internal node: byte 21 = NodeRayMask
// "ignored", plus the two spare bits of K0's pad byte 17; instance leaf: bytes 8-9 of the unused StartNodeAddress; TLAS header:
// bytes 32-33); child i then sits at P + (present children before i), an instance leaf counting three 64-byte units (leaf +
// BLAS header).  k_tb_layout verifies this for every treelet and marks the ones where it does not hold (node lists de-duplicated
// across shared BLASes) as not stageable.
#include "vsrt_device.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr uint32_t INST_NONE = 0x7FFFFFu;
constexpr uint32_t RAY_DEFERRED = 0xFFFFFFFFu;
constexpr int TB_THREADS = 128;
constexpr int TB_CUR_N = 32;                 // entries of `current` a lane keeps during a round (one per node with pending children)
constexpr uint32_t TB_SMEM_UNITS = 768;      // staging buffer: 48 KB = the reference's default max_treelet_size
constexpr uint32_t TB_NO_P = 0x3FFu;         // "no staged position" in the 10-bit field of an entry
constexpr uint32_t TB_DEAD = 0xFFFFFFFFu;    // key of a ray that needs no further round
#ifndef VSRT_TB_STAGE_MIN_RAYS
#define VSRT_TB_STAGE_MIN_RAYS 16
#endif
#ifndef VSRT_TB_STAGE_MIN_UNITS
#define VSRT_TB_STAGE_MIN_UNITS 4
#endif

struct TbRay {                // 64 bytes per ray, global memory
  float min_thit, min_thit_object;
  uint32_t closest_leaf, closest_inst;
  uint32_t cnt, ray_nodes;    // records emitted | total_nodes_accessed (bits 0..19) + procedural visits (bits 20..31)
  uint32_t cur_tid;           // current_treelet_root as a treelet index (tid_known) or the slot of the self-rooted node moved over from `other`
  uint32_t oth_n;
  uint32_t flags;             // bit 0 tid_known, bit 1 seed valid, bits 8..15 max level seen
  uint32_t pad[3];
  uint4 seed;                 // the TLAS root's entry when it lies in the TLAS header's treelet (consumed by the first round)
};
static_assert(sizeof(TbRay) == 64, "TbRay layout");

struct TbLayout {             // device tables of the treelet-layout copy
  const uint8_t* bytes;       // Σ units x 64 bytes
  const uint32_t* unit_off;   // [n_treelets + 1] first unit of every treelet
  const uint32_t* tflags;     // [n_treelets] bit 0: positions verified, the treelet may be staged
};

struct TbParams {
  TraverseParams tp;
  TbLayout lay;
  TbRay* state;               // [n_rays]
  uint4* ostack;              // [n_rays * stack_n] `other` LIFO of every ray
  uint32_t stack_n;
  const uint32_t* keys; const uint32_t* ids;   // sorted (treelet, ray) pairs of this round
  uint32_t* keys_next; uint32_t* ids_next;     // what every ray is filed under for the next round (same positions)
  uint32_t n_live;
  unsigned int* n_live_next;  // rays that need another round
  uint32_t free_run;          // != 0: the last launch -- no binning, no staging, every lane keeps switching treelets until its ray ends
  unsigned long long* stats;  // [0] rays processed, [1] rays served from a staged treelet, [2] CTAs that staged, [3] bytes staged, [4] node visits from shared memory, [5] node visits from the arena
};

VS_DEV uint32_t u_level(uint32_t meta) { return (meta >> 23) & 0xffu; }
VS_DEV uint32_t u_inst(uint32_t meta) { return meta & INST_NONE; }

VS_DEV Node64 load_node_smem(const uint4* sbuf, uint32_t pos) {
  const uint4* p = sbuf + (size_t)pos * 4u;
  Node64 n; const uint4 a = p[0], b = p[1], c = p[2], d = p[3];
  n.w[0] = a.x; n.w[1] = a.y; n.w[2] = a.z; n.w[3] = a.w; n.w[4] = b.x; n.w[5] = b.y; n.w[6] = b.z; n.w[7] = b.w;
  n.w[8] = c.x; n.w[9] = c.y; n.w[10] = c.z; n.w[11] = c.w; n.w[12] = d.x; n.w[13] = d.y; n.w[14] = d.z; n.w[15] = d.w;
  return n;
}

struct ActiveRay { Ray8 ray; Idir idir; float tmult; uint32_t inst; bool nonfinite; };

// ---------------------------------------------------------------- treelet-layout copy (built once per formation)
__global__ void k_tb_units(const unsigned long long* __restrict__ tl_off, const uint64_t* __restrict__ tl_node, uint32_t n_t, uint32_t* __restrict__ units) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n_t) return;
  uint32_t u = 0;
  for (unsigned long long k = tl_off[t]; k < tl_off[t + 1]; k++) u += ((uint32_t)(tl_node[k] >> 32) == K_INSTANCE) ? 2u : 1u;
  units[t] = u;
}
__global__ void k_tb_narrow(const unsigned long long* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = (uint32_t)in[i];
}
__global__ void k_tb_layout(const ArenaView av, const uint32_t* __restrict__ node_tid, const unsigned long long* __restrict__ tl_off, const uint64_t* __restrict__ tl_node,
                            const uint32_t* __restrict__ unit_off, uint32_t n_t, uint8_t* __restrict__ layout, uint32_t* __restrict__ tflags) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n_t) return;
  const unsigned long long k0 = tl_off[t]; const uint32_t n = (uint32_t)(tl_off[t + 1] - k0);
  uint8_t* dst0 = layout + (uint64_t)unit_off[t] * 64u;
  const uint32_t total = unit_off[t + 1] - unit_off[t];
  bool ok = total <= 1022u && n > 0;
#define ENT(i_) tl_node[k0 + (i_)]
#define UNITS_OF(e_) (((uint32_t)((e_) >> 32) == K_INSTANCE) ? 2u : 1u)
  uint32_t next = 1, posn = n ? UNITS_OF(ENT(0)) : 0u, posk = 0;
#define ADVANCE() do { posn += UNITS_OF(ENT(next)); next++; } while (0)
  if (n > 1 && (uint32_t)(ENT(0) >> 32) == K_INSTANCE) ADVANCE();      // a treelet rooted at an instance leaf: its BLAS header entry follows it
  for (uint32_t k = 0; k < n; k++) {
    const uint64_t ent = ENT(k); const uint32_t slot = (uint32_t)ent, kind = (uint32_t)(ent >> 32), un = UNITS_OF(ent);
    uint8_t* dst = dst0 + (uint64_t)posk * 64u;
    const uint4* src = reinterpret_cast<const uint4*>(av.base + (uint64_t)slot * 64u);
    for (uint32_t q = 0; q < un * 4u; q++) reinterpret_cast<uint4*>(dst)[q] = src[q];
    if (kind == K_TLAS_HEADER) {
      uint32_t rs = 0, P = 0xFFFFu;
      if (header_root(av, slot, rs) && next < n && (uint32_t)ENT(next) == rs) { P = posn; ADVANCE(); }
      dst[32] = (uint8_t)P; dst[33] = (uint8_t)(P >> 8);
    } else if (kind == K_TLAS_INTERNAL || kind == K_BLAS_INTERNAL) {
      if ((node_tid[slot] & VSRT_TID_MASK) != t) ok = false;                 // listed here, mapped to a higher treelet (shared BLAS)
      const uint32_t mask = dst[17] & 0x3fu;
      uint32_t child = slot + *reinterpret_cast<const uint32_t*>(dst + 12), P = TB_NO_P; bool gap = false;
      for (int i = 0; i < 6; i++) {
        const uint32_t sz = dst[22 + i] & 3u; if (!sz) continue;
        if ((mask >> i) & 1u) {
          if (gap || next >= n || (uint32_t)ENT(next) != child) { ok = false; break; }
          if (P == TB_NO_P) P = posn;
          const uint32_t ck = (uint32_t)(ENT(next) >> 32);
          ADVANCE();
          if (ck == K_INSTANCE) { if (next < n && (uint32_t)(ENT(next) >> 32) == K_BLAS_HEADER) ADVANCE(); else { ok = false; break; } }
        } else gap = true;                                                    // in-treelet children must be a prefix of the present ones
        child += sz;
      }
      if (P != TB_NO_P && P > 0x3FEu) { ok = false; P = TB_NO_P; }
      dst[21] = (uint8_t)P; dst[17] = (uint8_t)(mask | ((P >> 8) << 6));
    } else if (kind == K_INSTANCE) {
      uint32_t P = 0xFFFFu, rs = 0;
      if (k + 1 < n && (uint32_t)(ENT(k + 1) >> 32) == K_BLAS_HEADER && header_root(av, (uint32_t)ENT(k + 1), rs)) {
        if ((node_tid[rs] & VSRT_TID_MASK) == t) { if (next < n && (uint32_t)ENT(next) == rs) { P = posn; ADVANCE(); } else ok = false; }
      } else ok = false;
      dst[8] = (uint8_t)P; dst[9] = (uint8_t)(P >> 8);
    }
    posk += un;
  }
  if (next != n) ok = false;
  tflags[t] = ok ? 1u : 0u;
#undef ADVANCE
#undef UNITS_OF
#undef ENT
}

// ---------------------------------------------------------------- per-ray helpers shared by the init and the round kernel
struct Acc { unsigned int nodes, max_nodes, max_level, hits, term, rays, err; };

VS_DEV void tb_finalize(const TraverseParams& p, uint32_t r, const TbRay& s, uint32_t flags, float w_tmax, unsigned int* s_cnt) {
  const ArenaView& av = p.av; const uint8_t* base = av.base;
  const uint32_t cap = p.cap;
  uint32_t err = 0;
  if (s.cnt + (s.ray_nodes >> 20) > cap) err |= EF_TRACE_CAP;
  p.counts[r] = s.cnt; p.nproc[r] = s.ray_nodes >> 20;
  const uint32_t ray_nodes = s.ray_nodes & 0xFFFFFu;
  vsrt_hit h;
  h.hit_geometry = 0; h.world_min_thit = 0.0f; h.primitive_index = 0; h.geometry_index = 0; h.instance_index = 0;
  h.barycentric[0] = h.barycentric[1] = h.barycentric[2] = 0.0f;
  h.intersection_point[0] = h.intersection_point[1] = h.intersection_point[2] = 0.0f;
  h.n_all_hits = 0; h.instance_leaf_address = 0;
  if (s.min_thit < w_tmax) {
    atomicAdd(&s_cnt[3], 1u);
    const vsrt_ray* rp = p.rays + r;
    Ray8 w; w.ox = __ldg(&rp->origin[0]); w.oy = __ldg(&rp->origin[1]); w.oz = __ldg(&rp->origin[2]); w.tmin = __ldg(&rp->tmin);
    w.dx = __ldg(&rp->direction[0]); w.dy = __ldg(&rp->direction[1]); w.dz = __ldg(&rp->direction[2]); w.tmax = w_tmax;
    const Node64 q = load_node(base, s.closest_leaf);
    Ray8 o = w;
    const uint32_t inst_base = av.inst_base;
    if (s.closest_inst != INST_NONE) { InstCtx c; make_object_ray(base, inst_base + s.closest_inst, w, c); o = c.ray; }
    const uint32_t ci = inst_base + s.closest_inst;
    h.hit_geometry = 1; h.world_min_thit = s.min_thit;
    h.geometry_index = q.w[1] & 0x0fffffffu; h.primitive_index = q.w[2];
    h.instance_index = __ldg(reinterpret_cast<const uint32_t*>(base + (uint64_t)ci * 64u + 72u));
    h.intersection_point[0] = fadd(w.ox, fmul(w.dx, s.min_thit)); h.intersection_point[1] = fadd(w.oy, fmul(w.dy, s.min_thit)); h.intersection_point[2] = fadd(w.oz, fmul(w.dz, s.min_thit));
    barycentric(q, fadd(o.ox, fmul(o.dx, s.min_thit_object)), fadd(o.oy, fmul(o.dy, s.min_thit_object)), fadd(o.oz, fmul(o.dz, s.min_thit_object)), h.barycentric);
    h.instance_leaf_address = slot_to_host(av, ci);
  }
  p.hits[r] = h;
  atomicAdd(&s_cnt[0], ray_nodes); atomicMax(&s_cnt[1], ray_nodes); atomicMax(&s_cnt[2], (s.flags >> 8) & 0xffu);
  if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) atomicAdd(&s_cnt[5], 1u);
  atomicAdd(&s_cnt[6], 1u);
  if (err) atomicOr(&s_cnt[7], err);
}

VS_DEV void tb_flush_counters(const TraverseParams& p, unsigned int* s_cnt) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* c = p.counters->v;
    if (s_cnt[0]) atomicAdd(c + CI_TOT_NODES, (unsigned long long)s_cnt[0]);
    if (s_cnt[3]) atomicAdd(c + CI_NUM_HITS, (unsigned long long)s_cnt[3]);
    if (s_cnt[5]) atomicAdd(c + CI_N_ANYHIT_RAYS, (unsigned long long)s_cnt[5]);
    if (s_cnt[6] - s_cnt[5]) atomicAdd(c + CI_N_CLOSEST_RAYS, (unsigned long long)(s_cnt[6] - s_cnt[5]));
    if (s_cnt[6]) atomicAdd(c + CI_RAY_COUNT, (unsigned long long)s_cnt[6]);
    atomicMax(c + CI_MAX_NODES, (unsigned long long)s_cnt[1]);
    atomicMax(c + CI_MAX_DEPTH, (unsigned long long)s_cnt[2]);
    if (s_cnt[7]) atomicOr(p.err_flags, s_cnt[7]);
  }
}

// treelet a ray is filed under: the one that holds the node it takes from the front of `other` next
VS_DEV uint32_t tb_next_key(const TbParams& q, uint32_t r, uint32_t oth_n) {
  const uint4 t = q.ostack[(uint64_t)r * q.stack_n + (oth_n - 1u)];
  uint32_t ci; asm("bfind.u32 %0, %1;" : "=r"(ci) : "r"(t.z & 0x3F0000u));
  const uint32_t cb = __byte_perm(t.y, t.z, 0x7770u + (ci - 16u));
  const uint32_t tid = __ldg(q.tp.tv.node_tid + t.x + (cb & 15u));
  return tid == VSRT_NO_TID ? q.tp.tv.n_treelets : (tid & VSRT_TID_MASK);      // n_treelets = "in no treelet": sorts behind every real bin, never staged
}

// ---------------------------------------------------------------- round 0 set-up: one thread per ray (:1650-1741)
__global__ void __launch_bounds__(TB_THREADS) k_tb_init(const TbParams q) {
  const TraverseParams& p = q.tp; const ArenaView& av = p.av;
  __shared__ unsigned int s_cnt[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t r64 = (uint64_t)blockIdx.x * TB_THREADS + threadIdx.x;
  if (r64 < p.n_rays) {
    const uint32_t r = (uint32_t)r64;
    const vsrt_ray* rp = p.rays + r;
    Ray8 w; w.ox = __ldg(&rp->origin[0]); w.oy = __ldg(&rp->origin[1]); w.oz = __ldg(&rp->origin[2]); w.tmin = __ldg(&rp->tmin);
    w.dx = __ldg(&rp->direction[0]); w.dy = __ldg(&rp->direction[1]); w.dz = __ldg(&rp->direction[2]); w.tmax = __ldg(&rp->tmax);
    const uint32_t flags = __ldg(&rp->ray_flags);
    uint32_t key = TB_DEAD;
    if (ray_needs_exact(w) || av.force_exact) { p.counts[r] = RAY_DEFERRED; atomicOr(&s_cnt[7], (unsigned int)EF_NEED_EXACT); }
    else {
      TbRay s;
      s.min_thit = w.tmax; s.min_thit_object = 0.0f; s.closest_leaf = 0; s.closest_inst = INST_NONE; s.cnt = 0; s.ray_nodes = 0;
      s.cur_tid = VSRT_NO_TID; s.oth_n = 0; s.flags = 1u; s.pad[0] = s.pad[1] = s.pad[2] = 0; s.seed = make_uint4(0, 0, 0, 0);
      uint32_t* rstage = p.stage + (uint64_t)r * p.cap;
      if (s.cnt < p.cap) rstage[s.cnt] = (av.tlas_slot << 3) | (uint32_t)C_STRUCT;                  // :1685
      s.cnt++;
      uint32_t top_root = 0;
      if (!header_root(av, av.tlas_slot, top_root)) atomicOr(&s_cnt[7], (unsigned int)EF_BAD_BVH);
      else {
        const float* hb = reinterpret_cast<const float*>(av.base + (uint64_t)av.tlas_slot * 64u + 8u);
        float th; const Idir id = calc_idir(w);
        if (ray_box(__ldg(hb), __ldg(hb + 1), __ldg(hb + 2), __ldg(hb + 3), __ldg(hb + 4), __ldg(hb + 5), id, w, th)) {   // :1712-1741
          s.cur_tid = root_rank(p.tv, av.tlas_slot);                                                 // current_treelet_root = TLAS, :1707
          const uint32_t tr = __ldg(p.tv.node_tid + top_root);
          uint4 c = make_uint4(top_root, 0x10u, (1u << 16) | (TB_NO_P << 22), (1u << 23) | INST_NONE);   // one present child at offset 0
          s.flags |= 1u << 8;
          if ((tr & VSRT_TID_MASK) == s.cur_tid) {
            // staged position of the TLAS root inside the header's treelet (layout copy, bytes 32-33 of the header)
            if (s.cur_tid != VSRT_NO_TID && (__ldg(q.lay.tflags + s.cur_tid) & 1u)) {
              const uint8_t* hd = q.lay.bytes + (uint64_t)__ldg(q.lay.unit_off + s.cur_tid) * 64u;
              const uint32_t P = (uint32_t)hd[32] | ((uint32_t)hd[33] << 8);
              if (P < TB_NO_P) c.z = (1u << 16) | (P << 22);
            }
            s.seed = c; s.flags |= 2u; key = s.cur_tid;
          } else {
            if (tr & VSRT_TID_SELF_ROOTED) c.y |= 0x80u;
            q.ostack[(uint64_t)r * q.stack_n] = c; s.oth_n = 1;
            key = tr == VSRT_NO_TID ? p.tv.n_treelets : (tr & VSRT_TID_MASK);
          }
        }
      }
      if (key == TB_DEAD) tb_finalize(p, r, s, flags, w.tmax, s_cnt);     // missed the scene box: one record, no hit
      else q.state[r] = s;
    }
    q.keys_next[r] = key; q.ids_next[r] = r;
    if (key != TB_DEAD) atomicAdd(q.n_live_next, 1u);
  }
  tb_flush_counters(p, s_cnt);
}

// ---------------------------------------------------------------- one round
__global__ void __launch_bounds__(TB_THREADS) k_tb_round(const TbParams q) {
  const TraverseParams& p = q.tp; const ArenaView& av = p.av;
  const uint8_t* __restrict__ base = av.base;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  extern __shared__ __align__(128) uint4 s_nodes[];          // TB_SMEM_UNITS x 64 bytes
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ unsigned int s_cnt[8];
  __shared__ uint32_t s_stage_tid, s_stage_units;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  const uint32_t i = blockIdx.x * TB_THREADS + threadIdx.x;
  const bool live = i < q.n_live;
  const uint32_t key = live ? q.keys[i] : TB_DEAD;
  const uint32_t r = live ? q.ids[i] : 0u;

  // ---- which treelet does this CTA stage?  The one its middle ray is in, if enough of its rays share it.
  const bool free_run = q.free_run != 0u;
  const uint32_t mid = min(blockIdx.x * TB_THREADS + TB_THREADS / 2, q.n_live - 1u);
  const uint32_t cand = q.keys[mid];
  const int sharers = __syncthreads_count(live && key == cand);
  if (threadIdx.x == 0) {
    uint32_t units = 0;
    if (!free_run && cand < p.tv.n_treelets && sharers >= VSRT_TB_STAGE_MIN_RAYS && (__ldg(q.lay.tflags + cand) & 1u)) {
      units = __ldg(q.lay.unit_off + cand + 1u) - __ldg(q.lay.unit_off + cand);
      if (units < VSRT_TB_STAGE_MIN_UNITS || units > TB_SMEM_UNITS) units = 0;
    }
    s_stage_tid = units ? cand : TB_DEAD; s_stage_units = units;
    if (units) {
      const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const uint32_t stage_tid = s_stage_tid;
  if (stage_tid != TB_DEAD) {
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) {
      // ONE bulk copy brings the whole treelet: its nodes are contiguous in the treelet-layout copy
      const uint32_t bytes = s_stage_units * 64u;
      const uint8_t* src = q.lay.bytes + (uint64_t)__ldg(q.lay.unit_off + stage_tid) * 64u;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_nodes);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
      atomicAdd(q.stats + 2, 1ull); atomicAdd(q.stats + 3, (unsigned long long)bytes);
    }
    // every thread waits for the bytes (phase 0 of the barrier)
    asm volatile("{\n .reg .pred p;\n TB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra TB_WAIT;\n}" ::"r"(mb), "r"(0) : "memory");
  }
  // ---- lanes of one treelet find each other: the lowest lane of every group counts the group (regroup bookkeeping)
  const unsigned grp = __match_any_sync(full, key);
  const bool staged_lane = live && key == stage_tid;
  if (live && lane == __ffs(grp) - 1) { atomicAdd(q.stats + 0, (unsigned long long)__popc(grp)); if (staged_lane) atomicAdd(q.stats + 1, (unsigned long long)__popc(grp)); }

  uint32_t next_key = TB_DEAD;
  if (live) {
    TbRay s = q.state[r];
    const vsrt_ray* rp = p.rays + r;
    const uint32_t flags = __ldg(&rp->ray_flags);
    const float w_tmin = __ldg(&rp->tmin), w_tmax = __ldg(&rp->tmax);
    uint32_t* rstage = p.stage + (uint64_t)r * p.cap;
    uint4* ost = q.ostack + (uint64_t)r * q.stack_n;
    const uint32_t cap = p.cap, inst_base = av.inst_base;
    uint32_t cnt = s.cnt, ray_nodes = s.ray_nodes, oth_n = s.oth_n, cur_tid = s.cur_tid, max_level = (s.flags >> 8) & 0xffu, err = 0;
    bool tid_known = (s.flags & 1u) != 0u;
    float min_thit = s.min_thit, min_thit_object = s.min_thit_object;
    uint32_t closest_leaf = s.closest_leaf, closest_inst = s.closest_inst;
    uint4 cur[TB_CUR_N]; int cur_n = 0;
    bool took_other = (s.flags & 2u) != 0u, dead = false, deferred = false;   // the seed round stays in the TLAS header's treelet
    unsigned long long n_smem = 0, n_glob = 0;
    ActiveRay a; a.inst = 0xFFFFFFFFu; a.tmult = 1.0f; a.nonfinite = false;
    a.ray.ox = a.ray.oy = a.ray.oz = a.ray.dx = a.ray.dy = a.ray.dz = a.ray.tmin = a.ray.tmax = 0.0f; a.idir.x = a.idir.y = a.idir.z = 0.0f;
    if (s.flags & 2u) { cur[0] = s.seed; cur_n = 1; }

#define EMIT(slot_, code_) do { if (cnt < cap) rstage[cnt] = ((slot_) << 3) | (uint32_t)(code_); cnt++; } while (0)
#define CUR_TID() (tid_known ? cur_tid : (tid_known = true, cur_tid = __ldg(p.tv.node_tid + cur_tid) & VSRT_TID_MASK))
#define LOAD_WORLD(w_) do { (w_).ox = __ldg(&rp->origin[0]); (w_).oy = __ldg(&rp->origin[1]); (w_).oz = __ldg(&rp->origin[2]); (w_).tmin = w_tmin; \
      (w_).dx = __ldg(&rp->direction[0]); (w_).dy = __ldg(&rp->direction[1]); (w_).dz = __ldg(&rp->direction[2]); (w_).tmax = w_tmax; } while (0)
#define ACTIVATE(inst_) do { const uint32_t i_ = (inst_); if (a.inst != i_) { a.inst = i_; Ray8 w_; LOAD_WORLD(w_); \
      if (i_ == INST_NONE) { a.ray = w_; a.idir = calc_idir(w_); a.tmult = 1.0f; a.nonfinite = false; } \
      else { InstCtx c_; make_object_ray(base, inst_base + i_, w_, c_); a.ray = c_.ray; a.idir = c_.idir; a.tmult = c_.tmult; a.nonfinite = c_.exact; } } } while (0)

    while (!dead && !deferred) {
      // ================= take the next entry: from `current`, else -- once per round -- the front of `other` (:1748-1754)
      uint32_t eslot, emeta, spos = TB_NO_P; bool in_cur, is_leaf;
      if (cur_n) {
        const uint4 t = cur[cur_n - 1];
        uint32_t ci; asm("bfind.u32 %0, %1;" : "=r"(ci) : "r"(t.z & 0x3F0000u));
        const uint32_t z2 = t.z ^ (1u << ci);
        if (z2 & 0x3F0000u) cur[cur_n - 1].z = z2; else cur_n--;
        ci -= 16u;
        const uint32_t cb = __byte_perm(t.y, t.z, 0x7770u + ci);
        eslot = t.x + (cb & 15u); emeta = t.w; in_cur = true; is_leaf = (cb & 0x40u) != 0u;
        const uint32_t P = t.z >> 22;
        if (P != TB_NO_P && staged_lane) {
          // position of child ci: P + the present children before it (an instance leaf of a TLAS node takes leaf + BLAS header = 3 units)
          const bool top = u_inst(t.w) == INST_NONE;
          uint32_t pos = P;
          for (uint32_t j = 0; j < ci; j++) { const uint32_t bj = __byte_perm(t.y, t.z, 0x7770u + j); if (bj & 0x10u) pos += (top && (bj & 0x40u)) ? 3u : 1u; }
          spos = pos;
        }
      } else if ((!took_other || free_run) && oth_n) {
        took_other = true;
        const uint4 t = ost[oth_n - 1u];
        uint32_t ci; asm("bfind.u32 %0, %1;" : "=r"(ci) : "r"(t.z & 0x3F0000u));
        const uint32_t z2 = t.z ^ (1u << ci);
        if (z2 & 0x3F0000u) ost[oth_n - 1u].z = z2; else oth_n--;
        ci -= 16u;
        const uint32_t cb = __byte_perm(t.y, t.z, 0x7770u + ci);
        eslot = t.x + (cb & 15u); emeta = t.w; is_leaf = (cb & 0x40u) != 0u;
        const bool selfroot = (cb & 0x80u) != 0u;
        if (av.tlas_delta == 0) {
          in_cur = selfroot; cur_tid = eslot; tid_known = false;
          if (!selfroot) { cur_tid = root_rank(p.tv, eslot); tid_known = true; }
        } else { uint32_t s2; in_cur = false; tid_known = true; cur_tid = host_to_slot(av, slot_to_host(av, eslot) - (uint64_t)av.tlas_delta, s2) ? root_rank(p.tv, s2) : VSRT_NO_TID; }
        if (selfroot && in_cur && staged_lane) spos = 0;      // the root of a treelet is the first node of its list
      } else break;

      const uint32_t inst = u_inst(emeta);
      if (!is_leaf) {
        // ================= internal node (TLAS :1759-1875, BLAS :1954-2072)
        const Node64 n = spos != TB_NO_P ? load_node_smem(s_nodes, spos) : load_node(base, eslot);
        if (spos != TB_NO_P) n_smem++; else n_glob++;
        EMIT(eslot, inst == INST_NONE ? C_INTERNAL_TLAS : C_INTERNAL_BLAS); ray_nodes++;
        ACTIVATE(inst);
        if (a.nonfinite) { deferred = true; break; }
        const uint32_t mask = test_children<false>(n, a.ray, a.idir, fmul(min_thit, a.tmult), p.magic16);
        const uint32_t lo4 = __byte_perm(n.w[5], n.w[6], 0x5432), hi2 = n.w[6] >> 16;
        const uint32_t pre4 = (lo4 & 0x03030303u) * 0x01010101u;
        const uint32_t t4 = pre4 >> 24, xlo = pre4 << 8, xhi = t4 | ((t4 + (hi2 & 3u)) << 8);
        const uint32_t child0 = eslot + (uint32_t)node_child_offset(n);
        const uint32_t level = u_level(emeta), clevel = level < 255u ? level + 1u : 255u;
        if (mask && clevel > max_level) max_level = clevel;
        const uint32_t cmeta = (clevel << 23) | inst;
        uint32_t mc = node_byte(n, 17) & 0x3fu;
        if (!in_cur) {
          mc = 0;
          const uint32_t ct = CUR_TID();
          for (uint32_t m = mask; m; m &= m - 1u) {
            const uint32_t bit = m & (0u - m); uint32_t bi; asm("bfind.u32 %0, %1;" : "=r"(bi) : "r"(bit));
            if ((__ldg(p.tv.node_tid + child0 + __byte_perm(xlo, xhi, 0x7770u + bi)) & VSRT_TID_MASK) == ct) mc |= bit;
          }
        }
        const uint32_t mcur = mask & mc, moth = mask ^ mcur;
        // present bit (0x10) of every child byte: size != 0
        const uint32_t pr4 = ((lo4 | (lo4 >> 1)) & 0x01010101u) << 4, pr2 = ((hi2 | (hi2 >> 1)) & 0x0101u) << 4;
        const uint32_t ey = xlo | (lo4 & 0xC0C0C0C0u) | pr4, ez = xhi | (hi2 & 0xC0C0u) | pr2;
        if (moth) { if (oth_n >= q.stack_n) err |= EF_STACK; else { ost[oth_n] = make_uint4(child0, ey, ez | (moth << 16) | (TB_NO_P << 22), cmeta); oth_n++; } }
        if (mcur) {
          const uint32_t P = spos != TB_NO_P ? (node_byte(n, 21) | ((node_byte(n, 17) >> 6) << 8)) : TB_NO_P;
          if (cur_n >= TB_CUR_N) err |= EF_STACK; else { cur[cur_n] = make_uint4(child0, ey, ez | (mcur << 16) | (P << 22), cmeta); cur_n++; }
        }
      } else if (inst == INST_NONE) {
        // ================= instance leaf (:1876-1953)
        EMIT(eslot, C_INSTANCE); ray_nodes++;
        uint32_t hdr = 0, broot = 0;
        const uint32_t iref = eslot - inst_base;
        if (!instance_blas_header(av, eslot, hdr) || !header_root(av, hdr, broot) || eslot < inst_base || iref >= INST_NONE) { err |= EF_BAD_BVH; dead = true; break; }
        EMIT(hdr, C_STRUCT);
        uint4 c = make_uint4(broot, 0x10u, (1u << 16) | (TB_NO_P << 22), (u_level(emeta) << 23) | iref);
        const uint32_t tb = __ldg(p.tv.node_tid + broot);
        if ((tb & VSRT_TID_MASK) == CUR_TID()) {
          if (spos != TB_NO_P) {   // staged leaf: position of the BLAS root inside this treelet (bytes 8-9 of the leaf's layout copy)
            const uint32_t w2 = s_nodes[(size_t)spos * 4u].z, P = w2 & 0xFFFFu;
            if (P < TB_NO_P) c.z = (1u << 16) | (P << 22);
          }
          if (cur_n >= TB_CUR_N) err |= EF_STACK; else { cur[cur_n] = c; cur_n++; }
        } else {
          if (tb & VSRT_TID_SELF_ROOTED) c.y |= 0x80u;
          if (oth_n >= q.stack_n) err |= EF_STACK; else { ost[oth_n] = c; oth_n++; }
        }
      } else {
        // ================= BLAS leaf (:2073-2204)
        const Node64 ql = spos != TB_NO_P ? load_node_smem(s_nodes, spos) : load_node(base, eslot);
        if (spos != TB_NO_P) n_smem++; else n_glob++;
        EMIT(eslot, C_DESC);
        if (((ql.w[1] >> 29) & 1u) == 0u) {
          ACTIVATE(inst);
          if (a.nonfinite) { deferred = true; break; }
          float thit = 0.0f;
          const bool hit = ray_tri(ql, a.ray, thit);
          const float tw = !hit ? 0.0f : (a.tmult == 1.0f ? thit : fdiv(thit, a.tmult));
          const bool acc = hit && w_tmin <= tw && tw <= w_tmax && tw < min_thit;            // :2124-2127
          if (acc) {
            min_thit = tw; min_thit_object = thit; closest_leaf = eslot; closest_inst = inst;
            EMIT(eslot, C_QUAD_HIT); ray_nodes++;
            if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) { cur_n = 0; oth_n = 0; }      // :2151-2155
          } else { EMIT(eslot, C_QUAD); ray_nodes++; }
        } else {
          EMIT(eslot, C_PROC); ray_nodes++;
          const uint32_t j = ray_nodes >> 20;
          if (cnt + j < cap) rstage[cap - 1u - j] = inst;
          if (j < 0xFFFu && (ray_nodes & 0xFFFFFu) != 0xFFFFFu) ray_nodes += 1u << 20; else err |= EF_UNSUPPORTED;
        }
      }
    }
#undef ACTIVATE
#undef LOAD_WORLD
#undef CUR_TID
#undef EMIT
    if (n_smem) atomicAdd(q.stats + 4, n_smem);
    if (n_glob) atomicAdd(q.stats + 5, n_glob);
    if (err) atomicOr(&s_cnt[7], err);
    if (deferred) { p.counts[r] = RAY_DEFERRED; atomicOr(&s_cnt[7], (unsigned int)EF_NEED_EXACT); }
    else {
      s.min_thit = min_thit; s.min_thit_object = min_thit_object; s.closest_leaf = closest_leaf; s.closest_inst = closest_inst;
      s.cnt = cnt; s.ray_nodes = ray_nodes; s.cur_tid = cur_tid; s.oth_n = oth_n; s.flags = (tid_known ? 1u : 0u) | (max_level << 8);
      if (dead || oth_n == 0u) tb_finalize(p, r, s, flags, w_tmax, s_cnt);
      else { q.state[r] = s; next_key = tb_next_key(q, r, oth_n); }
    }
  }
  if (i < q.n_live) { q.keys_next[i] = next_key; q.ids_next[i] = r; }
  const unsigned more = __ballot_sync(full, next_key != TB_DEAD);
  if (lane == 0 && more) atomicAdd(q.n_live_next, (unsigned int)__popc(more));
  tb_flush_counters(p, s_cnt);
}

}  // namespace

// ================================================================= host side
struct TbTables { uint8_t* bytes; uint32_t* unit_off; uint32_t* tflags; uint64_t total_units; };

int vsrt_tb_build_layout(const ArenaView& av, const FormOutputs& fo, uint32_t n_treelets, void** tables_out, cudaStream_t st) {
  TbTables* T = new TbTables(); *T = TbTables{};
  uint32_t* units = nullptr; unsigned long long* off64 = nullptr; void* tmp = nullptr;
  unsigned long long total = 0;
  bool ok = cudaMalloc(&units, (size_t)std::max(n_treelets, 1u) * 4) == cudaSuccess && cudaMalloc(&off64, ((size_t)n_treelets + 1) * 8) == cudaSuccess &&
            cudaMalloc(&tmp, vsrt_scan_tmp_bytes(n_treelets)) == cudaSuccess && cudaMalloc(&T->unit_off, ((size_t)n_treelets + 1) * 4) == cudaSuccess &&
            cudaMalloc(&T->tflags, (size_t)std::max(n_treelets, 1u) * 4) == cudaSuccess;
  if (ok && n_treelets) {
    k_tb_units<<<(n_treelets + 255) / 256, 256, 0, st>>>((const unsigned long long*)fo.tl_off, fo.tl_node, n_treelets, units);
    ok = vsrt_launch_scan(units, n_treelets, (uint64_t*)off64, tmp, st) == VSRT_OK;
    k_tb_narrow<<<(n_treelets + 1 + 255) / 256, 256, 0, st>>>(off64, n_treelets + 1, T->unit_off);
    ok = ok && cudaMemcpyAsync(&total, off64 + n_treelets, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    ok = ok && total < (1ull << 32) && cudaMalloc(&T->bytes, (size_t)std::max<unsigned long long>(total, 1) * 64) == cudaSuccess;
    if (ok) {
      k_tb_layout<<<(n_treelets + 127) / 128, 128, 0, st>>>(av, fo.node_tid, (const unsigned long long*)fo.tl_off, fo.tl_node, T->unit_off, n_treelets, T->bytes, T->tflags);
      ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    }
  }
  cudaFree(units); cudaFree(off64); cudaFree(tmp);
  T->total_units = total;
  if (!ok) { cudaFree(T->bytes); cudaFree(T->unit_off); cudaFree(T->tflags); delete T; *tables_out = nullptr; return VSRT_E_CUDA; }
  *tables_out = T;
  return VSRT_OK;
}

void vsrt_tb_free_layout(void* tables) {
  TbTables* T = (TbTables*)tables; if (!T) return;
  cudaFree(T->bytes); cudaFree(T->unit_off); cudaFree(T->tflags); delete T;
}

size_t vsrt_tb_scratch_bytes(uint64_t n_rays, uint32_t stack_n) {
  const uint64_t np = (n_rays + 63) & ~63ull;
  return n_rays * sizeof(TbRay) + n_rays * (uint64_t)stack_n * 16 + 4 * np * 4 + vsrt_radix_tmp_bytes(n_rays) + 1024;
}

// Runs the whole batch in rounds.  stats_out (host, may be NULL): [0..5] as TbParams::stats, [6] rounds.
int vsrt_launch_traverse_tb(const TraverseParams& tp, void* tables, uint32_t stack_n, void* scratch, unsigned long long* stats_out, cudaStream_t st) {
  TbTables* T = (TbTables*)tables;
  const uint64_t n = tp.n_rays;
  if (n == 0) return VSRT_OK;
  const uint64_t np = (n + 63) & ~63ull;
  uint8_t* p = (uint8_t*)scratch;
  unsigned long long* d_stats = (unsigned long long*)p; unsigned int* d_live = (unsigned int*)(p + 64); p += 1024;
  TbRay* state = (TbRay*)p; p += n * sizeof(TbRay);
  uint4* ostack = (uint4*)p; p += n * (uint64_t)stack_n * 16;
  uint32_t* kA = (uint32_t*)p; p += np * 4; uint32_t* iA = (uint32_t*)p; p += np * 4;
  uint32_t* kB = (uint32_t*)p; p += np * 4; uint32_t* iB = (uint32_t*)p; p += np * 4;
  void* radix_tmp = p;
  if (cudaMemsetAsync(d_stats, 0, 128, st) != cudaSuccess) return VSRT_E_CUDA;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_tb_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TB_SMEM_UNITS * 64)); attr_set = true; }
  TbParams q; q.tp = tp; q.lay.bytes = T->bytes; q.lay.unit_off = T->unit_off; q.lay.tflags = T->tflags;
  q.state = state; q.ostack = ostack; q.stack_n = stack_n; q.stats = d_stats; q.n_live_next = d_live;
  q.keys = nullptr; q.ids = nullptr; q.keys_next = kA; q.ids_next = iA; q.n_live = (uint32_t)n;
  k_tb_init<<<(unsigned)((n + TB_THREADS - 1) / TB_THREADS), TB_THREADS, 0, st>>>(q);
  // bits of the sort key: treelet indices 0 .. n_treelets (the last = "in no treelet"); a finished ray's key is all ones
  int passes = 1; while (passes < 4 && (1ull << (8 * passes)) <= (unsigned long long)tp.tv.n_treelets + 1ull) passes++;
  // Binned rounds pay while many rays share treelets, i.e. for the first few treelet levels below the root; the tail -- a few
  // rays each in a treelet of its own, hundreds of rounds deep -- would be nothing but launch and sort overhead, so after
  // VSRT_TB_ROUNDS binned rounds (default 8; 0 = bin to the end) one last launch lets every remaining ray run to its end.
  unsigned long long max_rounds = 8;
  if (const char* e = getenv("VSRT_TB_ROUNDS")) max_rounds = (unsigned long long)atoll(e);
  q.free_run = 0;
  unsigned int live = 0; unsigned long long rounds = 0;
  if (cudaMemcpyAsync(&live, d_live, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return VSRT_E_CUDA;
  while (live) {
    // (treelet, ray) pairs of the live rays sit in the first `live` positions... after the first round they are scattered among
    // dead ones in [0, n_prev): sorting puts the dead (key 0xFFFFFFFF -> low 24/32 bits all ones) last
    const uint64_t n_sort = q.n_live;
    uint32_t* ks = nullptr; uint32_t* is = nullptr;
    q.free_run = (max_rounds && rounds >= max_rounds) ? 1u : 0u;
    // (the free-running launch still needs the live rays in front of the dead ones: same sort)
    int rc = vsrt_launch_radix_sort(kA, iA, kB, iB, n_sort, passes, radix_tmp, nullptr, &ks, &is, st); if (rc) return rc;
    if (cudaMemsetAsync(d_live, 0, 4, st) != cudaSuccess) return VSRT_E_CUDA;
    q.keys = ks; q.ids = is; q.n_live = live;
    // the round writes the next keys into the buffers the sort did NOT end in
    q.keys_next = (ks == kA) ? kB : kA; q.ids_next = (is == iA) ? iB : iA;
    k_tb_round<<<(live + TB_THREADS - 1) / TB_THREADS, TB_THREADS, TB_SMEM_UNITS * 64, st>>>(q);
    if (cudaGetLastError() != cudaSuccess) return VSRT_E_CUDA;
    // next round sorts from (keys_next, ids_next): make them the "A" pair
    if (q.keys_next != kA) { std::swap(kA, kB); std::swap(iA, iB); }
    if (cudaMemcpyAsync(&live, d_live, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return VSRT_E_CUDA;
    rounds++;
    if (rounds > 100000) return VSRT_E_CUDA;
  }
  if (stats_out) {
    if (cudaMemcpyAsync(stats_out, d_stats, 48, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return VSRT_E_CUDA;
    stats_out[6] = rounds;
  }
  return VSRT_OK;
}
