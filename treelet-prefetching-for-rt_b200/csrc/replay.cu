// replay.cu -- the consumers of the trace inside the RT unit (gpgpu-sim/shader.cc), as batched kernels over the trace of
// the last batch:
//   * rt_unit::sort_mem_accesses (:3012-3089): per-ray reordering of the access list by treelet, -sort_method 0 / 1;
//   * the treelet-prefetch vote of rt_unit::cycle (:3419-3560): the rays resident in a unit vote with the treelet of their
//     pending access, heuristics 0-3 decide what part of the winner's node list is prefetched;
//   * the 32-byte chunks that decision queues (:3566-3620), metadata rows included.
// Both are pure functions of (trace, treelet tables); the unit's history (last_prefetched_treelet, queue occupancy,
// prefetch_delay) stays with the caller's timing model.
#include "vsrt_device.cuh"
#include <algorithm>

namespace {

// ------------------------------------------------------------------ inverted treelet lists (for -sort_method 0)
// inv_off[slot] .. inv_off[slot + 1]: the (treelet, position in that treelet's node list) pairs that list the slot.
__global__ void k_inv_count(const uint64_t* __restrict__ tl_node, unsigned long long n_entries, uint32_t* __restrict__ cnt) {
  const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_entries) atomicAdd(cnt + (uint32_t)tl_node[k], 1u);
}
__global__ void k_inv_fill(const unsigned long long* __restrict__ tl_off, const uint64_t* __restrict__ tl_node, uint32_t n_treelets,
                           const unsigned long long* __restrict__ inv_off, uint32_t* __restrict__ cursor, uint2* __restrict__ inv) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_treelets) return;
  const unsigned long long k0 = tl_off[t], k1 = tl_off[t + 1];
  for (unsigned long long k = k0; k < k1; k++) {
    const uint32_t slot = (uint32_t)tl_node[k];
    inv[inv_off[slot] + atomicAdd(cursor + slot, 1u)] = make_uint2(t, (uint32_t)(k - k0));
  }
}

// ------------------------------------------------------------------ sort_mem_accesses
constexpr int SORT_WARPS = 4;
constexpr uint32_t SORT_SMEM_N = 256;    // records of a ray staged in shared memory; longer rays work on global memory

// One warp per ray.  Every record gets a 64-bit key (treelet's first-appearance index, position inside the treelet,
// original index); its output position is the number of smaller keys (n is a few dozen: counting beats sorting).
//   method 1: key = (first i with tag_i == tag, 0, index); the record written is the FIRST one with the same address (:3069).
//   method 0: key = (first-appearance index of the earliest visited treelet that lists the node, position in that treelet's
//             list, index): the strict order walks treelets in order of first appearance and, inside each, its node list (:3046).
template <int METHOD>
__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_trace(const unsigned long long* __restrict__ offsets, uint64_t n_rays,
                                                                const vsrt_txn* __restrict__ txns, const uint32_t* __restrict__ tids,
                                                                const uint32_t* __restrict__ stage, uint32_t cap,
                                                                const unsigned long long* __restrict__ inv_off, const uint2* __restrict__ inv,
                                                                vsrt_txn* __restrict__ out_txns, uint32_t* __restrict__ out_tids,
                                                                unsigned long long* __restrict__ key_scratch) {
  __shared__ unsigned long long s_addr[SORT_WARPS][SORT_SMEM_N];
  __shared__ unsigned long long s_key[SORT_WARPS][SORT_SMEM_N];
  __shared__ uint32_t s_tag[SORT_WARPS][SORT_SMEM_N];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t r = (uint64_t)blockIdx.x * SORT_WARPS + w;
  if (r >= n_rays) return;
  const unsigned long long j0 = offsets[r], j1 = offsets[r + 1];
  const uint32_t n = (uint32_t)(j1 - j0);
  const bool in_smem = n <= SORT_SMEM_N;
  // views of the ray's records: staged copies when they fit, the global arrays otherwise
  const uint32_t* T = tids + j0;
  unsigned long long* K = key_scratch + j0;
  if (in_smem) {
    for (uint32_t i = lane; i < n; i += 32) { s_addr[w][i] = txns[j0 + i].address; s_tag[w][i] = tids[j0 + i]; }
    __syncwarp();
    T = s_tag[w]; K = s_key[w];
  }
#define ADDR(i_) (in_smem ? s_addr[w][(i_)] : txns[j0 + (i_)].address)
  for (uint32_t i = lane; i < n; i += 32) {
    unsigned long long key;
    if (METHOD == 1) {
      const uint32_t t = T[i];
      uint32_t f = 0; while (T[f] != t) f++;
      key = ((unsigned long long)f << 44) | i;
    } else {
      const uint32_t slot = stage[r * (uint64_t)cap + i] >> 3;
      uint32_t best_f = 0xFFFFFu, best_pos = 0xFFFFFFu;
      for (unsigned long long e = inv_off[slot]; e < inv_off[slot + 1]; e++) {
        const uint2 tp = inv[e];
        uint32_t f = 0; while (f < n && T[f] != tp.x) f++;          // is that treelet visited by this ray, and when first
        if (f < n && (f < best_f || (f == best_f && tp.y < best_pos))) { best_f = f; best_pos = tp.y; }
      }
      key = ((unsigned long long)best_f << 44) | ((unsigned long long)best_pos << 20) | i;   // unlisted nodes sort last, in order
    }
    K[i] = key;
  }
  __syncwarp();
  for (uint32_t i = lane; i < n; i += 32) {
    const unsigned long long key = K[i];
    uint32_t pos = 0;
    for (uint32_t j = 0; j < n; j++) pos += (K[j] < key) ? 1u : 0u;
    uint32_t src = i;
    if (METHOD == 1) { const unsigned long long a = ADDR(i); src = 0; while (ADDR(src) != a) src++; }
    out_txns[j0 + pos] = txns[j0 + src];
    out_tids[j0 + pos] = tids[j0 + src];
  }
#undef ADDR
}

// ------------------------------------------------------------------ prefetch vote
constexpr int VOTE_THREADS = 128;
constexpr uint32_t VOTE_SMEM_N = 4096;

struct VoteParams {
  const unsigned long long* offsets; const uint32_t* tids; uint64_t n_rays_batch;
  const unsigned long long* group_offsets; const unsigned long long* ray_ids; const uint32_t* front;
  uint64_t n_groups;
  const uint32_t* tl_root; const unsigned long long* tl_off; uint32_t n_treelets;
  uint32_t heuristic; double threshold;
  vsrt_prefetch_decision* out;     // root holds the treelet INDEX here; the host turns it into the root address
  uint32_t* hist_scratch;          // [n_treelets] zeros, used by groups larger than VOTE_SMEM_N (one such group per launch)
};

__device__ __forceinline__ uint32_t pending_tag(const VoteParams& p, unsigned long long pos_in_group) {
  const unsigned long long r = p.ray_ids ? p.ray_ids[pos_in_group] : pos_in_group;
  if (r >= p.n_rays_batch) return VSRT_NO_TID;
  const unsigned long long k = p.offsets[r] + (p.front ? p.front[r] : 0u);
  if (k >= p.offsets[r + 1]) return VSRT_NO_TID;                       // RT_mem_accesses.empty(): no vote (:3426)
  return p.tids[k];
}
__device__ __forceinline__ void decide(const VoteParams& p, uint64_t g, unsigned long long best, uint32_t total) {
  vsrt_prefetch_decision d;
  d.treelet_root = ~0ull; d.votes = 0; d.total = total; d.submit = 0; d.n_nodes = 0; d.first_node = 0; d.num_nodes = 0;
  if (best) {
    const uint32_t votes = (uint32_t)(best >> 32), t = ~(uint32_t)best;
    const uint32_t n_nodes = (uint32_t)(p.tl_off[t + 1] - p.tl_off[t]);
    const double pct = (double)votes / (double)total;                                     // :3496
    const uint32_t part = (uint32_t)(int)((double)n_nodes * pct + 0.5);                   // :3528
    d.treelet_root = t; d.votes = votes; d.n_nodes = n_nodes;
    d.submit = p.heuristic == 1 ? (pct >= p.threshold ? 1u : 0u) : 1u;
    d.num_nodes = (p.heuristic == 2 || p.heuristic == 3) ? part : n_nodes;
    d.first_node = p.heuristic == 3 ? n_nodes - part : 0u;
  }
  p.out[g] = d;
}
// candidate = votes << 32 | ~treelet index: the maximum is the most voted treelet, lowest index (= lowest root address,
// the std::map iteration order with strict '>', :3441) among equals
__device__ __forceinline__ unsigned long long block_max(unsigned long long v, unsigned long long* sh) {
  for (int o = 16; o; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, v, o); v = y > v ? y : v; }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long m = 0;
  for (int i = 0; i < VOTE_THREADS / 32; i++) m = sh[i] > m ? sh[i] : m;
  __syncthreads();
  return m;
}
__global__ void __launch_bounds__(VOTE_THREADS) k_vote_small(const VoteParams p) {
  __shared__ uint32_t s_tag[VOTE_SMEM_N];
  __shared__ unsigned long long s_red[VOTE_THREADS / 32];
  __shared__ unsigned int s_total;
  const uint64_t g = blockIdx.x;
  const unsigned long long g0 = p.group_offsets[g], g1 = p.group_offsets[g + 1];
  const uint32_t m = (uint32_t)(g1 - g0);
  if (m > VOTE_SMEM_N) return;                      // left to the large-group path
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (uint32_t i = threadIdx.x; i < m; i += VOTE_THREADS) { const uint32_t t = pending_tag(p, g0 + i); s_tag[i] = t; mine += t != VSRT_NO_TID; }
  if (mine) atomicAdd(&s_total, mine);
  __syncthreads();
  unsigned long long best = 0;
  for (uint32_t i = threadIdx.x; i < m; i += VOTE_THREADS) {
    const uint32_t t = s_tag[i];
    if (t == VSRT_NO_TID) continue;
    uint32_t c = 0;
    for (uint32_t j = 0; j < m; j++) c += s_tag[j] == t;
    const unsigned long long cand = ((unsigned long long)c << 32) | (uint32_t)~t;
    best = cand > best ? cand : best;
  }
  best = block_max(best, s_red);
  if (threadIdx.x == 0) decide(p, g, best, s_total);
}
// a group larger than VOTE_SMEM_N: histogram in global scratch, three passes over the group
__global__ void k_vote_large_count(const VoteParams p, uint64_t g) {
  const unsigned long long g0 = p.group_offsets[g], m = p.group_offsets[g + 1] - g0;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t t = pending_tag(p, g0 + i);
  if (t != VSRT_NO_TID) atomicAdd(p.hist_scratch + t, 1u);
}
__global__ void k_vote_large_pick(const VoteParams p, uint64_t g, unsigned long long* best, unsigned int* total) {
  const unsigned long long g0 = p.group_offsets[g], m = p.group_offsets[g + 1] - g0;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t t = pending_tag(p, g0 + i);
  if (t == VSRT_NO_TID) return;
  atomicAdd(total, 1u);
  atomicMax(best, ((unsigned long long)p.hist_scratch[t] << 32) | (uint32_t)~t);
}
__global__ void k_vote_large_finish(const VoteParams p, uint64_t g, unsigned long long* best, unsigned int* total) {
  const unsigned long long g0 = p.group_offsets[g], m = p.group_offsets[g + 1] - g0;
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) decide(p, g, *best, *total);
  if (i >= m) return;
  const uint32_t t = pending_tag(p, g0 + i);
  if (t != VSRT_NO_TID) p.hist_scratch[t] = 0;       // leave the scratch zeroed for the next large group
}

// ------------------------------------------------------------------ schedule_next_warp (:4307-4392)
// One warp of threads per RT unit; it walks the unit's resident warps in m_current_warps order, lane l looking at thread l.
__global__ void k_schedule_pick(const unsigned long long* __restrict__ offsets, const uint32_t* __restrict__ tids, uint64_t n_rays_batch,
                                const unsigned long long* __restrict__ unit_offsets, const unsigned long long* __restrict__ warp_ray_ids,
                                const uint8_t* __restrict__ stalled, const uint32_t* __restrict__ target_tid, const uint32_t* __restrict__ front,
                                uint64_t n_units, int scheduler, long long* __restrict__ pick) {
  const uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= n_units) return;
  const int lane = threadIdx.x & 31;
  const uint32_t target = target_tid[u];
  long long first_free = -1, best = -1; uint32_t best_n = 0;
  for (unsigned long long w = unit_offsets[u]; w < unit_offsets[u + 1]; w++) {
    if (stalled && stalled[w]) continue;
    if (first_free < 0) first_free = (long long)w;
    if ((scheduler != 1 && scheduler != 2) || target == VSRT_NO_TID) break;
    const unsigned long long r = warp_ray_ids[32ull * w + lane];
    bool match = false;
    if (r < n_rays_batch) {
      const unsigned long long k = offsets[r] + (front ? front[r] : 0u);
      match = k < offsets[r + 1] && tids[k] == target;
    }
    const uint32_t m = __popc(__ballot_sync(0xffffffffu, match));
    if (scheduler == 1 && m) { best = (long long)w; break; }
    if (scheduler == 2 && m > best_n) { best_n = m; best = (long long)w; }
  }
  if (lane == 0) pick[u] = best >= 0 ? best : first_free;
}

// ------------------------------------------------------------------ shader-table events (Baseline tables)
// One thread per ray walks the ray's staged records.  A procedural-leaf record is an intersection-table call (its instance
// is the j-th word from the end of the staging segment, left there by K1); in traceRay a QUAD_LEAF_HIT record of a ray
// without the Opaque flag is an any-hit call, whose instance is the last INSTANCE_LEAF record before it (traceRay drains a
// BLAS before it returns to the TLAS, :2679).  FILL = false counts, FILL = true writes the events and the any-hit Hit_data
// (same transform, triangle test and barycentrics as the traversal: make_object_ray / ray_tri / barycentric).
template <bool FILL>
__global__ void __launch_bounds__(128) k_table_events(const TableParams p) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.n_rays) return;
  const ArenaView& av = p.av;
  const uint32_t cnt = min(p.counts[r], p.cap);
  const uint32_t* seg = p.stage + r * (uint64_t)p.cap;
  const bool anyhit_ray = p.mode == VSRT_MODE_DFS && !(__ldg(&p.rays[r].ray_flags) & VSRT_RAY_FLAG_OPAQUE);
  const uint32_t inst_base = av.inst_base;   // lowest instance-leaf slot of the TLAS (K0)
  if (!FILL) {
    uint32_t n_any = 0;
    if (anyhit_ray) for (uint32_t k = 0; k < cnt; k++) n_any += (seg[k] & 7u) == C_QUAD_HIT;
    p.ev_counts[r] = p.nproc[r] + n_any;
    return;
  }
  const uint32_t tid = p.tid_x ? p.tid_x[r] : (uint32_t)(r & 31u);
  uint32_t row[2] = { 0u, 0u };
  if (p.tid_x) {   // threads of the CTA that share this tid_x and come earlier have advanced its row counters
    for (uint64_t q = r & ~31ull; q < r; q++) if (p.tid_x[q] == tid) { row[0] += p.nproc[q]; row[1] += p.ev_counts[q] - p.nproc[q]; }
  }
  unsigned long long o = p.ev_offsets[r];
  uint32_t last_inst = 0xFFFFFFFFu, j = 0;
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t rec = seg[k], slot = rec >> 3, code = rec & 7u;
    if (code == C_INSTANCE) { last_inst = slot; continue; }
    int table = -1; uint32_t inst_slot = 0;
    if (code == C_PROC) { table = 0; inst_slot = inst_base + seg[p.cap - 1u - j]; j++; }
    else if (code == C_QUAD_HIT && anyhit_ray && last_inst != 0xFFFFFFFFu) { table = 1; inst_slot = last_inst; }
    if (table < 0) continue;
    if (o < p.capacity) {
      const uint8_t* il = av.base + (uint64_t)inst_slot * 64u;
      const uint8_t* leaf = av.base + (uint64_t)slot * 64u;
      vsrt_table_event e;
      e.table = (uint32_t)table; e.shader_counter = row[table]; e.tid = tid; e.record = k; e.reserved = 0;
      e.hit_group_index = __ldg(reinterpret_cast<const uint32_t*>(il + 4)) & 0x00ffffffu;
      e.instance_id = __ldg(reinterpret_cast<const uint32_t*>(il + 72));
      e.primitive_id = __ldg(reinterpret_cast<const uint32_t*>(leaf + (table ? 8 : 12)));
      p.events[o] = e;
      if (p.anyhit) {
        vsrt_hit h;
        h.hit_geometry = 0; h.world_min_thit = 0.0f; h.primitive_index = 0; h.geometry_index = 0; h.instance_index = 0;
        h.barycentric[0] = h.barycentric[1] = h.barycentric[2] = 0.0f;
        h.intersection_point[0] = h.intersection_point[1] = h.intersection_point[2] = 0.0f;
        h.n_all_hits = 0; h.instance_leaf_address = 0;
        if (table == 1) {
          const vsrt_ray* rp = p.rays + r;
          Ray8 w;
          w.ox = __ldg(&rp->origin[0]); w.oy = __ldg(&rp->origin[1]); w.oz = __ldg(&rp->origin[2]); w.tmin = __ldg(&rp->tmin);
          w.dx = __ldg(&rp->direction[0]); w.dy = __ldg(&rp->direction[1]); w.dz = __ldg(&rp->direction[2]); w.tmax = __ldg(&rp->tmax);
          InstCtx c; make_object_ray(av.base, inst_slot, w, c);
          const Node64 q = load_node(av.base, slot);
          float thit = 0.0f;
          ray_tri(q, c.ray, thit);
          const float tw = fdiv(thit, c.tmult);                                     // anyhit_thit, :2895
          h.hit_geometry = 1; h.world_min_thit = tw;
          h.geometry_index = q.w[1] & 0x0fffffffu; h.primitive_index = q.w[2]; h.instance_index = e.instance_id;
          h.intersection_point[0] = fadd(w.ox, fmul(w.dx, tw)); h.intersection_point[1] = fadd(w.oy, fmul(w.dy, tw)); h.intersection_point[2] = fadd(w.oz, fmul(w.dz, tw));
          barycentric(q, fadd(c.ray.ox, fmul(c.ray.dx, thit)), fadd(c.ray.oy, fmul(c.ray.dy, thit)), fadd(c.ray.oz, fmul(c.ray.dz, thit)), h.barycentric);
          h.instance_leaf_address = slot_to_host(av, inst_slot);
        }
        p.anyhit[o] = h;
      }
    }
    row[table]++; o++;
  }
}

// ------------------------------------------------------------------ prefetch chunks
struct ChunkParams {
  ArenaView av; TreeletView tv;
  const unsigned long long* tl_off; const uint64_t* tl_node; const uint64_t* remap;
  const vsrt_prefetch_decision* dec; uint64_t n_groups;   // dec[].treelet_root = treelet INDEX (device copy)
  uint32_t load_metadata; uint32_t per_meta; unsigned long long metadata_base;
  const unsigned long long* chunk_off; unsigned long long* chunk_addr; unsigned long long* chunk_owner; unsigned long long capacity;
  uint32_t* counts;
};
__device__ __forceinline__ unsigned long long entry_address(const ChunkParams& p, uint64_t e) {
  const uint32_t slot = (uint32_t)e, kind = (uint32_t)(e >> 32);
  if (p.remap) return p.remap[slot];
  int64_t delta = p.av.tlas_delta;
  if (kind == K_BLAS_HEADER) { int64_t d; if (blas_delta_of(p.av, slot, d)) delta = d; }       // keyed by its allocBLAS address, :1149-1153
  return slot_to_host(p.av, slot) + (uint64_t)delta;
}
// one thread per group; FILL = false counts, FILL = true writes (address, owner) pairs at chunk_off[g]
template <bool FILL>
__global__ void k_prefetch_chunks(const ChunkParams p) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_groups) return;
  const vsrt_prefetch_decision d = p.dec[g];
  unsigned long long n = 0, o = FILL ? p.chunk_off[g] : 0;
  if (d.submit && d.treelet_root != ~0ull) {
    const unsigned long long k0 = p.tl_off[(uint32_t)d.treelet_root];
    for (uint32_t j = d.first_node; j < d.first_node + d.num_nodes && j < d.n_nodes; j++) {
      const uint64_t e = p.tl_node[k0 + j];
      const uint32_t size = ((uint32_t)(e >> 32) == K_INSTANCE) ? 128u : 64u;
      if (p.load_metadata) {
        // treelet_addr_to_metadata_idx[NODE address] (:3571): the node's own index if it is a root, else operator[] gives 0
        const uint32_t rk = root_rank(p.tv, (uint32_t)e);
        const unsigned long long ma = p.metadata_base + (unsigned long long)(rk == VSRT_NO_TID ? 0u : rk) * p.per_meta;
        for (uint32_t q = 0; q < p.per_meta / 32u; q++) { if (FILL && o + n < p.capacity) { p.chunk_addr[o + n] = ma + q * 32ull; p.chunk_owner[o + n] = ma; } n++; }
      }
      const unsigned long long a = FILL ? entry_address(p, e) : 0ull;
      for (uint32_t q = 0; q < (size + 31u) / 32u; q++) { if (FILL && o + n < p.capacity) { p.chunk_addr[o + n] = a + q * 32ull; p.chunk_owner[o + n] = a; } n++; }
    }
  }
  if (!FILL) p.counts[g] = (uint32_t)n;
}

}  // namespace

int vsrt_launch_build_inverse(const FormOutputs& fo, uint32_t n_treelets, uint64_t n_entries, uint32_t n_slots,
                              uint64_t** inv_off_out, uint2** inv_out, cudaStream_t st) {
  uint32_t* cnt = nullptr; uint32_t* cursor = nullptr; unsigned long long* off = nullptr; uint2* inv = nullptr; void* tmp = nullptr;
  int rc = VSRT_E_CUDA;
  if (cudaMalloc(&cnt, (size_t)(n_slots + 1) * 4) != cudaSuccess) goto done;
  if (cudaMalloc(&cursor, (size_t)(n_slots + 1) * 4) != cudaSuccess) goto done;
  if (cudaMalloc(&off, (size_t)(n_slots + 2) * 8) != cudaSuccess) goto done;
  if (cudaMalloc(&inv, (size_t)std::max<uint64_t>(n_entries, 1) * sizeof(uint2)) != cudaSuccess) goto done;
  if (cudaMalloc(&tmp, vsrt_scan_tmp_bytes(n_slots)) != cudaSuccess) goto done;
  cudaMemsetAsync(cnt, 0, (size_t)(n_slots + 1) * 4, st); cudaMemsetAsync(cursor, 0, (size_t)(n_slots + 1) * 4, st);
  if (n_entries) k_inv_count<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(fo.tl_node, n_entries, cnt);
  if ((rc = vsrt_launch_scan(cnt, n_slots, (uint64_t*)off, tmp, st)) != VSRT_OK) goto done;
  rc = VSRT_E_CUDA;
  if (n_treelets) k_inv_fill<<<(n_treelets + 127) / 128, 128, 0, st>>>((const unsigned long long*)fo.tl_off, fo.tl_node, n_treelets, off, cursor, inv);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) goto done;
  *inv_off_out = (uint64_t*)off; *inv_out = inv; off = nullptr; inv = nullptr;
  rc = VSRT_OK;
done:
  cudaFree(cnt); cudaFree(cursor); cudaFree(off); cudaFree(inv); cudaFree(tmp);
  return rc;
}

int vsrt_launch_sort_trace(int method, const uint64_t* offsets, uint64_t n_rays, const vsrt_txn* txns, const uint32_t* tids,
                           const uint32_t* stage, uint32_t cap, const uint64_t* inv_off, const uint2* inv,
                           vsrt_txn* out_txns, uint32_t* out_tids, uint64_t* key_scratch, cudaStream_t st) {
  if (n_rays == 0) return VSRT_OK;
  const unsigned grid = (unsigned)((n_rays + SORT_WARPS - 1) / SORT_WARPS);
  if (method == 1)
    k_sort_trace<1><<<grid, SORT_WARPS * 32, 0, st>>>((const unsigned long long*)offsets, n_rays, txns, tids, stage, cap, nullptr, nullptr, out_txns, out_tids, (unsigned long long*)key_scratch);
  else
    k_sort_trace<0><<<grid, SORT_WARPS * 32, 0, st>>>((const unsigned long long*)offsets, n_rays, txns, tids, stage, cap, (const unsigned long long*)inv_off, inv, out_txns, out_tids, (unsigned long long*)key_scratch);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_prefetch_vote(const uint64_t* offsets, const uint32_t* tids, uint64_t n_rays_batch, const uint64_t* group_offsets_dev,
                              const uint64_t* group_offsets_host, const uint64_t* ray_ids_dev, const uint32_t* front_dev, uint64_t n_groups,
                              const FormOutputs& fo, uint32_t n_treelets, uint32_t heuristic, double threshold,
                              vsrt_prefetch_decision* out_dev, cudaStream_t st) {
  if (n_groups == 0) return VSRT_OK;
  VoteParams p;
  p.offsets = (const unsigned long long*)offsets; p.tids = tids; p.n_rays_batch = n_rays_batch;
  p.group_offsets = (const unsigned long long*)group_offsets_dev; p.ray_ids = (const unsigned long long*)ray_ids_dev; p.front = front_dev; p.n_groups = n_groups;
  p.tl_root = fo.tl_root; p.tl_off = (const unsigned long long*)fo.tl_off; p.n_treelets = n_treelets;
  p.heuristic = heuristic; p.threshold = threshold; p.out = out_dev; p.hist_scratch = nullptr;
  k_vote_small<<<(unsigned)n_groups, VOTE_THREADS, 0, st>>>(p);
  if (cudaGetLastError() != cudaSuccess) return VSRT_E_CUDA;
  // groups that do not fit the shared-memory path (a whole frame voted as one unit): one at a time through a global histogram
  unsigned long long* best = nullptr; unsigned int* total = nullptr; int rc = VSRT_OK;
  for (uint64_t g = 0; g < n_groups && rc == VSRT_OK; g++) {
    const uint64_t m = group_offsets_host[g + 1] - group_offsets_host[g];
    if (m <= VOTE_SMEM_N) continue;
    if (!p.hist_scratch) {
      if (cudaMalloc(&p.hist_scratch, (size_t)std::max(n_treelets, 1u) * 4) != cudaSuccess || cudaMalloc(&best, 8) != cudaSuccess || cudaMalloc(&total, 4) != cudaSuccess) { rc = VSRT_E_CUDA; break; }
      cudaMemsetAsync(p.hist_scratch, 0, (size_t)n_treelets * 4, st);
    }
    cudaMemsetAsync(best, 0, 8, st); cudaMemsetAsync(total, 0, 4, st);
    const unsigned grid = (unsigned)((m + 255) / 256);
    k_vote_large_count<<<grid, 256, 0, st>>>(p, g);
    k_vote_large_pick<<<grid, 256, 0, st>>>(p, g, best, total);
    k_vote_large_finish<<<grid, 256, 0, st>>>(p, g, best, total);
    if (cudaGetLastError() != cudaSuccess) rc = VSRT_E_CUDA;
  }
  if (p.hist_scratch) { cudaStreamSynchronize(st); cudaFree(p.hist_scratch); cudaFree(best); cudaFree(total); }
  return rc;
}

int vsrt_launch_prefetch_chunks(bool fill, const ArenaView& av, const TreeletView& tv, const FormOutputs& fo, const uint64_t* remap,
                                const vsrt_prefetch_decision* dec_dev, uint64_t n_groups, uint32_t load_metadata, uint32_t per_meta, uint64_t metadata_base,
                                uint32_t* counts, const uint64_t* chunk_off, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t capacity, cudaStream_t st) {
  if (n_groups == 0) return VSRT_OK;
  ChunkParams p;
  p.av = av; p.tv = tv; p.tl_off = (const unsigned long long*)fo.tl_off; p.tl_node = fo.tl_node; p.remap = remap;
  p.dec = dec_dev; p.n_groups = n_groups; p.load_metadata = load_metadata; p.per_meta = per_meta; p.metadata_base = metadata_base;
  p.chunk_off = (const unsigned long long*)chunk_off; p.chunk_addr = (unsigned long long*)chunk_addr; p.chunk_owner = (unsigned long long*)chunk_owner; p.capacity = capacity;
  p.counts = counts;
  const unsigned grid = (unsigned)((n_groups + 127) / 128);
  if (fill) k_prefetch_chunks<true><<<grid, 128, 0, st>>>(p); else k_prefetch_chunks<false><<<grid, 128, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_schedule_pick(const uint64_t* offsets, const uint32_t* tids, uint64_t n_rays_batch, const uint64_t* unit_offsets_dev,
                              const uint64_t* warp_ray_ids_dev, const uint8_t* stalled_dev, const uint32_t* target_tid_dev, const uint32_t* front_dev,
                              uint64_t n_units, int scheduler, int64_t* pick_dev, cudaStream_t st) {
  if (n_units == 0) return VSRT_OK;
  k_schedule_pick<<<(unsigned)((n_units + 3) / 4), 128, 0, st>>>((const unsigned long long*)offsets, tids, n_rays_batch, (const unsigned long long*)unit_offsets_dev,
                                                                 (const unsigned long long*)warp_ray_ids_dev, stalled_dev, target_tid_dev, front_dev, n_units, scheduler,
                                                                 (long long*)pick_dev);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

// ------------------------------------------------------------------ Function_Call_Coalescing intersection table
// Coalescing_warp_intersection_table::add_intersection (intersection_table.cc:43-98) replayed over the table-0 events of a
// batch.  The rows are shared by the 32 threads of a CTA and filled lane by lane (execute_warp_inst_t runs the lanes in
// order), so a CTA is a sequential chain: one WARP per CTA, the row scan of one call done by the lanes in parallel -- lane l
// holds rows l, l + 32, l + 64, l + 96 (key + thread mask) in registers, "first row of this hit group whose thread_mask[tid]
// is free" is a ballot per 32 rows.
__global__ void __launch_bounds__(128) k_coalescing(const unsigned long long* __restrict__ ev_off, const vsrt_table_event* __restrict__ ev,
                                                    uint64_t n_rays, vsrt_coalescing_event* __restrict__ out, uint32_t* __restrict__ err) {
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t r0 = g * 32u;
  if (r0 >= n_rays) return;
  const uint64_t r1 = min(r0 + 32u, n_rays);
  uint32_t key[4] = { 0, 0, 0, 0 }, mask[4] = { 0, 0, 0, 0 }, n_rows = 0;
  for (uint64_t r = r0; r < r1; r++) {
    uint32_t seen = 0;
    for (unsigned long long k = ev_off[r]; k < ev_off[r + 1]; k++) {
      vsrt_coalescing_event o; o.row = o.appended = o.n_loads = o.first_new_load = 0;
      if (ev[k].table == 0) {
        const uint32_t want = ev[k].hit_group_index, bit = 1u << (ev[k].tid & 31u);
        uint32_t row = 0xFFFFFFFFu;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const unsigned m = __ballot_sync(0xffffffffu, (uint32_t)c * 32u + lane < n_rows && key[c] == want && !(mask[c] & bit));
          if (m && row == 0xFFFFFFFFu) row = (uint32_t)c * 32u + (uint32_t)(__ffs(m) - 1);
        }
        if (row != 0xFFFFFFFFu) { o.row = row; o.n_loads = row + 1u; }
        else if (n_rows >= 100u) { if (lane == 0) atomicOr(err, (uint32_t)EF_UNSUPPORTED); return; }   // the reference's allocation (INTERSECTION_TABLE_MAX_LENGTH rows) ends here
        else { row = n_rows; o.row = row; o.appended = 1; o.n_loads = n_rows; n_rows++; }
#pragma unroll
        for (int c = 0; c < 4; c++) if (row == (uint32_t)c * 32u + lane) { if (o.appended) { key[c] = want; mask[c] = bit; } else mask[c] |= bit; }
        o.first_new_load = min(seen, o.n_loads);
        seen = max(seen, o.n_loads);
      }
      if (lane == 0) out[k] = o;
    }
  }
}

int vsrt_launch_coalescing(const uint64_t* ev_off, const vsrt_table_event* ev, uint64_t n_rays, vsrt_coalescing_event* out, uint32_t* err, cudaStream_t st) {
  if (n_rays == 0) return VSRT_OK;
  const uint64_t n_groups = (n_rays + 31) / 32;
  k_coalescing<<<(unsigned)((n_groups + 3) / 4), 128, 0, st>>>((const unsigned long long*)ev_off, ev, n_rays, out, err);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_table_events(bool fill, const TableParams& p, uint32_t*, cudaStream_t st) {
  if (p.n_rays == 0) return VSRT_OK;
  const unsigned grid = (unsigned)((p.n_rays + 127) / 128);
  if (fill) k_table_events<true><<<grid, 128, 0, st>>>(p); else k_table_events<false><<<grid, 128, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}
