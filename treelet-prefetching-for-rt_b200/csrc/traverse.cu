// traverse.cu -- K1: per-ray functional traversal, both reference variants.
//
//   MODE_DFS      VulkanRayTracing::traceRay            (vulkan_ray_tracing.cc:2309-3076)
//   MODE_TREELET  VulkanRayTracing::traceRayWithTreelets (vulkan_ray_tracing.cc:1522-2307)
//
// One thread owns one ray from the first record to the hit record, so the visit order is the reference's by
// construction: DFS = LIFO stack where the first box-hit internal child is followed immediately (:2573,:2760);
// TREELET = two LIFOs, children go to `current` when they live in the current treelet, else to `other`
// (:1832-1856), `other` is drained only when `current` is empty (:1748-1754).  The kernel emits COMPACT trace
// records (slot << 3 | code) into the ray's staging segment; scan + K3 (compact.cu) turn them into the reference's
// 16-byte MemoryTransactionRecords in CSR order.
#include "vsrt_device.cuh"

namespace {

struct Entry { uint32_t slot; uint32_t meta; };   // meta = leaf << 31 | instance-leaf slot (VSRT_NO_INST = top level)
VS_DEV bool e_leaf(const Entry& e) { return (e.meta >> 31) != 0; }
VS_DEV uint32_t e_inst(const Entry& e) { return e.meta & 0x7FFFFFFFu; }
VS_DEV bool e_top(const Entry& e) { return e_inst(e) == VSRT_NO_INST; }

template <int MODE, int STACK_N>
__global__ void __launch_bounds__(128) k_traverse(const TraverseParams p) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = r < p.n_rays;
  uint32_t total_nodes = 0, max_level = 0, n_hit = 0, n_any = 0, n_term = 0, err = 0;

  if (active) {
    const ArenaView& av = p.av;
    const uint8_t* __restrict__ base = av.base;
    const vsrt_ray* rp = p.rays + r;
    Ray8 w;
    w.ox = __ldg(&rp->origin[0]); w.oy = __ldg(&rp->origin[1]); w.oz = __ldg(&rp->origin[2]); w.tmin = __ldg(&rp->tmin);
    w.dx = __ldg(&rp->direction[0]); w.dy = __ldg(&rp->direction[1]); w.dz = __ldg(&rp->direction[2]); w.tmax = __ldg(&rp->tmax);
    const uint32_t flags = __ldg(&rp->ray_flags);
    const bool terminate = (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) != 0;   // :1650 / :2411
    const bool opaque = (flags & VSRT_RAY_FLAG_OPAQUE) != 0;                      // skipAnyHitShader, :2413
    n_term = terminate ? 1u : 0u;
    const Idir widir = calc_idir(w);

    uint32_t* __restrict__ out = p.stage + r * (uint64_t)p.cap;
    const uint32_t cap = p.cap;
    uint32_t cnt = 0;
#define EMIT(slot_, code_) do { if (cnt < cap) out[cnt] = ((slot_) << 3) | (uint32_t)(code_); cnt++; } while (0)

    Entry stk[STACK_N];
    uint8_t lvl[STACK_N];
    int cur_n = 0, oth_n = 0;               // DFS uses cur only; TREELET: cur grows up from 0, other down from the top
    uint32_t cur_tid = VSRT_NO_TID;

    float min_thit = w.tmax, min_thit_object = 0.0f;                             // :1671
    uint32_t closest_leaf = 0, closest_inst = VSRT_NO_INST, n_all_hits = 0;
    InstCtx ctx; ctx.inst_slot = VSRT_NO_INST; ctx.tmult = 1.0f; ctx.ray = w; ctx.idir = widir;

    EMIT(av.tlas_slot, C_STRUCT);                                                // :1685 / :2447
    uint32_t top_root = 0;
    bool have_next = false; Entry next; next.slot = 0; next.meta = 0; uint32_t next_lvl = 0;
    if (!header_root(av, av.tlas_slot, top_root)) err |= EF_BAD_BVH;
    else {
      // scene box of the TLAS header (:1712-1741 / :2472-2484)
      const float* hb = reinterpret_cast<const float*>(base + (uint64_t)av.tlas_slot * 64u + 8u);
      float th;
      if (ray_box(__ldg(hb), __ldg(hb + 1), __ldg(hb + 2), __ldg(hb + 3), __ldg(hb + 4), __ldg(hb + 5), widir, w, th)) {
        Entry e; e.slot = top_root; e.meta = VSRT_NO_INST;
        if (MODE == VSRT_MODE_TREELET) {
          cur_tid = root_rank(p.tv, av.tlas_slot);                               // current_treelet_root = TLAS + offset, :1707
          if (__ldg(p.tv.node_tid + top_root) == cur_tid) { stk[cur_n] = e; lvl[cur_n] = 1; cur_n++; }
          else { stk[STACK_N - 1] = e; lvl[STACK_N - 1] = 1; oth_n = 1; }
        } else { stk[0] = e; lvl[0] = 1; cur_n = 1; }
        max_level = 1;
      }
    }

    while (true) {
      Entry e; uint32_t level;
      if (MODE == VSRT_MODE_DFS && have_next) { e = next; level = next_lvl; have_next = false; }
      else {
        if (cur_n == 0) {
          if (MODE == VSRT_MODE_DFS || oth_n == 0) break;
          // :1748-1754 -- move the front of `other` into `current`; current_treelet_root becomes that node's HOST
          // address, which equals a treelet's device address only for the root at (host - tlas_delta).
          Entry m = stk[STACK_N - oth_n]; uint8_t ml = lvl[STACK_N - oth_n]; oth_n--;
          stk[0] = m; lvl[0] = ml; cur_n = 1;
          if (av.tlas_delta == 0) cur_tid = root_rank(p.tv, m.slot);
          else { uint32_t s2; cur_tid = host_to_slot(av, slot_to_host(av, m.slot) - (uint64_t)av.tlas_delta, s2) ? root_rank(p.tv, s2) : VSRT_NO_TID; }
        }
        cur_n--; e = stk[cur_n]; level = lvl[cur_n];
      }
      const bool top = e_top(e);
      if (!e_leaf(e)) {
        // ---- internal node (TLAS :1759-1875 / :2500-2599, BLAS :1954-2072 / :2687-2786)
        const Node64 n = load_node(base, e.slot);
        EMIT(e.slot, top ? C_INTERNAL_TLAS : C_INTERNAL_BLAS); total_nodes++;
        if (!top && ctx.inst_slot != e_inst(e)) make_object_ray(base, e_inst(e), w, ctx);
        const uint32_t mask = top ? test_children(n, w, widir, min_thit) : test_children(n, ctx.ray, ctx.idir, fmul(min_thit, ctx.tmult));
        uint32_t child = e.slot + (uint32_t)node_child_offset(n);
        const uint32_t clevel = level < 255u ? level + 1u : 255u;
        if (mask) { if (clevel > max_level) max_level = clevel; }
#pragma unroll
        for (int i = 0; i < 6; i++) {
          const uint32_t info = node_child_info(n, i);
          if ((mask >> i) & 1u) {
            Entry c; c.slot = child; c.meta = (e.meta & 0x7FFFFFFFu) | (((info >> 2) != 0u) ? 0x80000000u : 0u);
            if (MODE == VSRT_MODE_DFS) {
              if ((info >> 2) == 0u && !have_next) { next = c; next_lvl = clevel; have_next = true; }   // first hit internal child
              else if (cur_n < STACK_N) { stk[cur_n] = c; lvl[cur_n] = (uint8_t)clevel; cur_n++; }
              else err |= EF_STACK;
            } else {
              if (cur_n + oth_n >= STACK_N) err |= EF_STACK;
              else if (__ldg(p.tv.node_tid + child) == cur_tid) { stk[cur_n] = c; lvl[cur_n] = (uint8_t)clevel; cur_n++; }
              else { oth_n++; stk[STACK_N - oth_n] = c; lvl[STACK_N - oth_n] = (uint8_t)clevel; }
            }
          }
          child += info & 3u;
        }
      } else if (top) {
        // ---- instance leaf (:1876-1953 / :2602-2677)
        EMIT(e.slot, C_INSTANCE); total_nodes++;
        uint32_t hdr = 0, broot = 0;
        if (!instance_blas_header(av, e.slot, hdr) || !header_root(av, hdr, broot)) { err |= EF_BAD_BVH; break; }
        EMIT(hdr, C_STRUCT);                                                     // BLAS header record, :1913 / :2645
        make_object_ray(base, e.slot, w, ctx);
        Entry c; c.slot = broot; c.meta = e.slot;                                // BLAS root inherits the leaf's level (:1944)
        if (MODE == VSRT_MODE_DFS) { if (cur_n < STACK_N) { stk[cur_n] = c; lvl[cur_n] = (uint8_t)level; cur_n++; } else err |= EF_STACK; }
        else {
          if (cur_n + oth_n >= STACK_N) err |= EF_STACK;
          else if (__ldg(p.tv.node_tid + broot) == cur_tid) { stk[cur_n] = c; lvl[cur_n] = (uint8_t)level; cur_n++; }
          else { oth_n++; stk[STACK_N - oth_n] = c; lvl[STACK_N - oth_n] = (uint8_t)level; }
        }
      } else {
        // ---- BLAS leaf (:2073-2204 / :2789-2985)
        EMIT(e.slot, C_DESC);
        const Node64 q = load_node(base, e.slot);
        if (((q.w[1] >> 29) & 1u) == 0u) {
          if (ctx.inst_slot != e_inst(e)) make_object_ray(base, e_inst(e), w, ctx);
          float thit = 0.0f;
          const bool hit = ray_tri(q, ctx.ray, thit);
          const float tw = fdiv(thit, ctx.tmult);
          bool acc = hit && w.tmin <= tw && tw <= w.tmax;                         // :2843
          if (MODE == VSRT_MODE_TREELET) acc = acc && tw < min_thit;              // :2127
          if (acc) {
            if (MODE == VSRT_MODE_TREELET) min_thit = tw;
            else { if (opaque && tw < min_thit) min_thit = tw; if (!opaque) { n_all_hits++; n_any++; } }   // :2850, :2869-2929
            min_thit_object = thit; closest_leaf = e.slot; closest_inst = e_inst(e);
            EMIT(e.slot, C_QUAD_HIT); total_nodes++;
            if (terminate) { cur_n = 0; oth_n = 0; have_next = false; }           // :2151-2155 / :2932-2935
          } else { EMIT(e.slot, C_QUAD); total_nodes++; }
        } else { EMIT(e.slot, C_PROC); total_nodes++; }                           // intersection-table transactions: not built yet
      }
    }
#undef EMIT
    if (cnt > cap) err |= EF_TRACE_CAP;
    p.counts[r] = cnt;

    // ---- hit record (:2211-2245 / :2990-3033)
    vsrt_hit h;
    h.hit_geometry = 0; h.world_min_thit = 0.0f; h.primitive_index = 0; h.geometry_index = 0; h.instance_index = 0;
    h.barycentric[0] = h.barycentric[1] = h.barycentric[2] = 0.0f;
    h.intersection_point[0] = h.intersection_point[1] = h.intersection_point[2] = 0.0f;
    h.n_all_hits = n_all_hits; h.instance_leaf_address = 0;
    if (min_thit < w.tmax) {
      n_hit = 1;
      const Node64 q = load_node(base, closest_leaf);
      if (ctx.inst_slot != closest_inst) make_object_ray(base, closest_inst, w, ctx);
      h.hit_geometry = 1; h.world_min_thit = min_thit;
      h.geometry_index = q.w[1] & 0x0fffffffu; h.primitive_index = q.w[2];
      h.instance_index = __ldg(reinterpret_cast<const uint32_t*>(base + (uint64_t)closest_inst * 64u + 72u));
      h.intersection_point[0] = fadd(w.ox, fmul(w.dx, min_thit));
      h.intersection_point[1] = fadd(w.oy, fmul(w.dy, min_thit));
      h.intersection_point[2] = fadd(w.oz, fmul(w.dz, min_thit));
      barycentric(q, fadd(ctx.ray.ox, fmul(ctx.ray.dx, min_thit_object)), fadd(ctx.ray.oy, fmul(ctx.ray.dy, min_thit_object)),
                  fadd(ctx.ray.oz, fmul(ctx.ray.dz, min_thit_object)), h.barycentric);
      h.instance_leaf_address = slot_to_host(av, closest_inst);
    }
    p.hits[r] = h;
  }

  // ---- functional counters (cuda-sim.h:155-166): warp-reduce, one atomic per warp
  const unsigned full = 0xffffffffu;
  const uint32_t s_nodes = __reduce_add_sync(full, total_nodes), s_hit = __reduce_add_sync(full, n_hit);
  const uint32_t s_any = __reduce_add_sync(full, n_any), s_term = __reduce_add_sync(full, n_term), s_act = __reduce_add_sync(full, active ? 1u : 0u);
  const uint32_t m_nodes = __reduce_max_sync(full, total_nodes), m_lvl = __reduce_max_sync(full, max_level), e_all = __reduce_or_sync(full, err);
  if ((threadIdx.x & 31) == 0) {
    unsigned long long* c = p.counters->v;
    atomicAdd(c + CI_TOT_NODES, (unsigned long long)s_nodes);
    if (s_hit) atomicAdd(c + CI_NUM_HITS, (unsigned long long)s_hit);
    if (s_any) atomicAdd(c + CI_NUM_ANY_HITS, (unsigned long long)s_any);
    if (s_term) atomicAdd(c + CI_N_ANYHIT_RAYS, (unsigned long long)s_term);
    if (s_act - s_term) atomicAdd(c + CI_N_CLOSEST_RAYS, (unsigned long long)(s_act - s_term));
    atomicMax(c + CI_MAX_NODES, (unsigned long long)m_nodes);
    atomicMax(c + CI_MAX_DEPTH, (unsigned long long)m_lvl);
    if (e_all) atomicOr(p.err_flags, e_all);
  }
}

template <int STACK_N>
int launch_n(const TraverseParams& p, cudaStream_t st) {
  const unsigned block = 128;
  const uint64_t grid = (p.n_rays + block - 1) / block;
  if (grid == 0) return VSRT_OK;
  if (grid > 0x7fffffffull) return VSRT_E_INVALID;
  if (p.mode == VSRT_MODE_TREELET) k_traverse<VSRT_MODE_TREELET, STACK_N><<<(unsigned)grid, block, 0, st>>>(p);
  else k_traverse<VSRT_MODE_DFS, STACK_N><<<(unsigned)grid, block, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

}  // namespace

int vsrt_launch_traverse(const TraverseParams& p, uint32_t stack_entries, cudaStream_t st) {
  if (stack_entries <= 96) return launch_n<96>(p, st);
  if (stack_entries <= 192) return launch_n<192>(p, st);
  return launch_n<384>(p, st);
}
