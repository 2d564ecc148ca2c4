// compact.cu -- exclusive scan of per-ray record counts and K3, the trace-compaction kernel that expands the
// compact staging records into the reference's MemoryTransactionRecord {address,size,type}
// (abstract_hardware_model.h:315-321) in the reference's per-ray transaction order, CSR over rays, plus the
// treelet index of every record (addrToTreeletID, vulkan_ray_tracing.cc:468-472, as an index into the ascending
// root table == treelet_addr_to_metadata_idx).  Also accumulates g_rt_mem_access_type[], accessedDataSize
// (:2257-2261) and the per-treelet visit histogram the prefetcher's popularity vote is built from (shader.cc:3424-3433).
#include "vsrt_device.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 32;   // 8192 counts per tile: a 2 M-ray frame is 254 tiles, i.e. 8 look-back rounds for the last one (with 2048 per tile the look-back chain alone took 25 us)
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_scan_excl(unsigned long long v, unsigned long long* total, unsigned long long* sh /*[32]*/) {
  // inclusive warp scan
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  if (wid == 0) {
    unsigned long long s = lane < (SCAN_THREADS / 32) ? sh[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    sh[lane] = s;   // inclusive over warps
  }
  __syncthreads();
  const unsigned long long warp_off = wid ? sh[wid - 1] : 0ull;
  *total = sh[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return warp_off + x - v;
}

// One-pass exclusive scan (decoupled look-back): a block takes the next tile from a ticket counter -- so every tile before it
// has at least started --, publishes its tile total, then one warp walks back over the preceding tiles' status words until it
// meets one whose inclusive prefix is already known.  A status word carries its flag in the top two bits and the value in the
// other 62, so flag and value arrive together and no fence is needed.  One launch instead of three (5 + 8 + 12 us on the
// 2 M-ray frame); `tmp` = [ticket][status per tile], zeroed by the launcher.
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_VALUE = (1ull << 62) - 1ull;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(const uint32_t* __restrict__ counts, uint64_t n, unsigned long long* __restrict__ tmp,
                                                              unsigned long long* __restrict__ offsets, unsigned long long* __restrict__ total_out) {
  __shared__ unsigned long long sh[32];
  __shared__ unsigned long long s_tile, s_prefix;
  unsigned long long* const status = tmp + 1;
  if (threadIdx.x == 0) s_tile = atomicAdd(tmp, 1ull);
  __syncthreads();
  const uint64_t tile = s_tile;
  const uint64_t base = tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t c[SCAN_ITEMS]; unsigned long long s = 0;
  const bool vec = base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(counts) & 15u) == 0;     // the thread's 32 counts as eight 16-byte loads
  if (vec) {
    const uint4* q = reinterpret_cast<const uint4*>(counts + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; i++) { const uint4 v = __ldg(q + i); c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w; }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) c[i] = base + i < n ? counts[base + i] : 0u;
  }
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) s += c[i];
  unsigned long long tot;
  const unsigned long long ex = block_scan_excl(s, &tot, sh);
  if (threadIdx.x < 32) {
    unsigned long long prefix = 0;
    if (tile == 0) { if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(status) = ST_PREFIX | tot; }
    else {
      if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(status + tile) = ST_AGG | tot;
      // look back, 32 tiles at a time: lane l reads tile (end - 1 - l); everything after the nearest tile with a known prefix counts
      long long end = (long long)tile;
      for (;;) {
        const long long t = end - 1 - (long long)threadIdx.x;
        unsigned long long w = ST_PREFIX;                                  // tiles before 0: an empty prefix
        if (t >= 0) { do { w = *reinterpret_cast<const volatile unsigned long long*>(status + t); } while ((w >> 62) == 0ull); }
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
        const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;        // nearest tile (lowest lane) whose prefix is known
        unsigned long long v = ((int)threadIdx.x <= first && t >= 0) ? (w & ST_VALUE) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        prefix += v;
        if (has_prefix) break;
        end -= 32;
      }
      if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(status + tile) = ST_PREFIX | ((prefix + tot) & ST_VALUE);
    }
    if (threadIdx.x == 0) s_prefix = prefix;
  }
  __syncthreads();
  unsigned long long o = s_prefix + ex;
  if (vec && (reinterpret_cast<uintptr_t>(offsets) & 15u) == 0) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(offsets + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 2; i++) { ulonglong2 v; v.x = o; o += c[2 * i]; v.y = o; o += c[2 * i + 1]; q[i] = v; }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) offsets[base + i] = o; o += c[i]; }
  }
  if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) { offsets[n] = o; if (total_out) *total_out = o; }   // the thread that holds the last element also writes the grand total
}

// ---------------------------------------------------------------- K3
#ifndef VSRT_K3_THREADS
#define VSRT_K3_THREADS 128   // the warps of a CTA wait for each other before the table flush: 512 -> 0.93 ms, 256 -> 0.68, 128 -> 0.64, 64 -> 0.63, 32 -> 0.84
#endif
constexpr int K3_THREADS = VSRT_K3_THREADS;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int K3_RAYS = K3_WARPS * 32;   // rays per CTA: every warp owns 32 consecutive rays and their contiguous output range
#ifndef VSRT_K3_ILP
#define VSRT_K3_ILP 4
#endif
constexpr int K3_ILP = VSRT_K3_ILP;        // independent 32-record windows in flight per warp
#ifndef VSRT_K3_HASH_BITS
#define VSRT_K3_HASH_BITS 10
#endif
#ifndef VSRT_K3_PERSIST
#define VSRT_K3_PERSIST 0   // measured: a persistent grid saturates the CTA-private table and is slower (1.10 vs 0.72 ms)
#endif
#ifndef VSRT_K3_WARP_TABLE
#define VSRT_K3_WARP_TABLE 0   // A/B: 1 = every warp has its own slice of the table and flushes it itself (no CTA barrier at the end); slower, 0.77 vs 0.74 ms
#endif
constexpr int K3_HASH_BITS = VSRT_K3_HASH_BITS; // CTA-private treelet histogram: 2^bits (key,count) slots in shared memory, flushed once per CTA
constexpr int K3_CTAS_PER_SM = 4; // only for the persistent-grid A/B variant (VSRT_K3_PERSIST=1)

__device__ __forceinline__ uint32_t code_size(uint32_t code) { return code == C_INSTANCE ? 128u : (code == C_DESC ? 8u : 64u); }
__device__ __forceinline__ uint32_t code_type(uint32_t code) { return code == C_INTERNAL_TLAS ? (uint32_t)VSRT_TXN_BVH_INTERNAL_NODE : code; }

// One warp expands the records of 32 consecutive rays.  The warp's output range [offsets[r], offsets[r + 32]) is
// contiguous; it is walked in windows of 32 records, one record per lane, so the 16-byte and 4-byte stores of a window are
// single coalesced transactions.  Which ray a record belongs to is found without a search: the lanes hold the 32 start
// offsets, one REDUX.OR builds the bitmap of ray starts inside the window and a popcount of the bits at or below the lane
// gives the ray (every ray has at least its TLAS-header record; a warp that sees an empty ray counts with shuffles instead).
#ifndef VSRT_K3_MIN_BLOCKS
#define VSRT_K3_MIN_BLOCKS 10   // x 128 threads = 1280 threads per SM at 51 registers, no spills (256 x 4: 0.71 ms, 256 x 5: 0.68 ms)
#endif
// SIMPLE = one host span, one host->device offset for every buffer, original addresses: a record's address is one multiply-add
// PACKED = the output is the 4-byte packed record (slot << 3 | code) per transaction instead of the 16-byte record + treelet index:
//          the form a host consumer takes over PCIe (vsrt_trace_rays_packed); counters and histogram are accumulated either way
template <bool SIMPLE, bool PACKED>
__global__ void __launch_bounds__(K3_THREADS, VSRT_K3_MIN_BLOCKS) k_compact(const CompactParams p) {
  // queued by the host before it knows whether the traversal succeeded and how many records there are (see run_batch)
  if (p.err_flags && (*reinterpret_cast<const volatile uint32_t*>(p.err_flags) & p.fatal_mask)) return;
  if (p.offsets[p.n_rays] > p.out_capacity) return;
  __shared__ unsigned int s_hist[8];
  __shared__ unsigned int s_hkey[1 << K3_HASH_BITS], s_hcnt[1 << K3_HASH_BITS];
  for (uint32_t i = threadIdx.x; i < (1u << K3_HASH_BITS); i += K3_THREADS) { s_hkey[i] = VSRT_NO_TID; s_hcnt[i] = 0; }
  if (threadIdx.x < 8) s_hist[threadIdx.x] = 0;
  __syncthreads();
  // the warp's slice of the table (VSRT_K3_WARP_TABLE) or the whole table
  constexpr int TBITS = VSRT_K3_WARP_TABLE ? K3_HASH_BITS - 3 : K3_HASH_BITS;
  unsigned int* const t_key = s_hkey + (VSRT_K3_WARP_TABLE ? (threadIdx.x >> 5) << TBITS : 0);
  unsigned int* const t_cnt = s_hcnt + (VSRT_K3_WARP_TABLE ? (threadIdx.x >> 5) << TBITS : 0);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const ArenaView& av = p.av;
  const bool one_span = av.n_spans == 1;
  const uint64_t span_host = one_span ? av.spans[0].host : 0ull;
  const uint64_t simple_base = span_host + (uint64_t)av.tlas_delta;
  uint32_t hc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  const uint64_t n_blocks = (p.n_rays + K3_RAYS - 1) / K3_RAYS;
  for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const uint64_t rw0 = blk * K3_RAYS + (uint64_t)(threadIdx.x >> 5) * 32u;
    if (rw0 >= p.n_rays) continue;
    const uint32_t nr = (uint32_t)min((uint64_t)32, p.n_rays - rw0);
    const unsigned long long my_off = p.offsets[rw0 + min((uint32_t)lane, nr)];
    const unsigned long long j0 = __shfl_sync(full, my_off, 0), j1 = p.offsets[rw0 + nr];
    const uint32_t total = (uint32_t)(j1 - j0);
    const bool rvalid = (uint32_t)lane < nr;
    const uint32_t rel = rvalid ? (uint32_t)(my_off - j0) : 0xFFFFFFFFu;
    const uint32_t nxt = __shfl_down_sync(full, rel, 1);
    const bool any_empty = __ballot_sync(full, rvalid && (((uint32_t)lane + 1u < nr ? nxt : total) == rel)) != 0u;
    const uint32_t* stage_w = p.stage + rw0 * (uint64_t)p.cap;
    uint32_t packed_lo = 0, packed_hi = 0, since_flush = 0;
    for (uint32_t wb = 0; wb < total; wb += 32u * K3_ILP) {
      uint32_t pos[K3_ILP], k[K3_ILP], ray[K3_ILP], rec[K3_ILP], tid[K3_ILP]; bool valid[K3_ILP];
#pragma unroll
      for (int u = 0; u < K3_ILP; u++) {
        const uint32_t wbase = wb + 32u * u;
        pos[u] = wbase + (uint32_t)lane; valid[u] = pos[u] < total;
        if (!any_empty) {
          const uint32_t nle = __popc(__ballot_sync(full, rel <= wbase));                     // rays that start at or before the window
          const uint32_t dw = rel - wbase;
          const uint32_t bit = (dw - 1u < 31u) ? (1u << dw) : 0u;                             // 1 <= rel - wbase <= 31
          const uint32_t starts = __reduce_or_sync(full, bit);                                  // ray starts inside the window
          ray[u] = nle - 1u + __popc(starts & ((2u << lane) - 1u));
        } else {
          uint32_t c = 0;
          for (int m = 0; m < 32; m++) { const uint32_t rm = __shfl_sync(full, rel, m); c += (rm <= pos[u]) ? 1u : 0u; }
          ray[u] = c - 1u;
        }
        if (!valid[u]) ray[u] = 0;
        k[u] = pos[u] - __shfl_sync(full, rel, (int)ray[u]);
      }
#pragma unroll
      for (int u = 0; u < K3_ILP; u++) rec[u] = valid[u] ? __ldg(stage_w + (ray[u] * p.cap + k[u])) : 0u;   // < 32 * cap: 32-bit index inside the warp's 32 segments
#pragma unroll
      for (int u = 0; u < K3_ILP; u++) { tid[u] = valid[u] ? __ldg(p.tv.node_tid + (rec[u] >> 3)) : VSRT_NO_TID; if (tid[u] != VSRT_NO_TID) tid[u] &= VSRT_TID_MASK; }
#pragma unroll
      for (int u = 0; u < K3_ILP; u++) {
        if (valid[u]) {
          const uint32_t slot = rec[u] >> 3, code = rec[u] & 7u;
          // host -> simulated-device offset the reference applies to this record
          int64_t delta = av.tlas_delta;
          if (!SIMPLE && !av.uniform_delta) {
            const uint32_t* seg = stage_w + (uint64_t)ray[u] * p.cap;
            if (code == C_STRUCT && k[u] > 0) { int64_t d; if (blas_delta_of(av, slot, d)) delta = d; }          // :1908-1913 / :2640-2645
            else if (p.mode == VSRT_MODE_DFS && code != C_INTERNAL_TLAS && code != C_INSTANCE && k[u] > 0) {
              // traceRay keeps device_offset = offset of the BLAS it is inside (:2640) until the next TLAS node (:2503,:2605)
              for (uint32_t b = k[u]; b-- > 0;) { const uint32_t pr = __ldg(seg + b); if ((pr & 7u) == C_STRUCT && b > 0) { int64_t d; if (blas_delta_of(av, pr >> 3, d)) delta = d; break; } }
            }
          }
          // original_bvh_to_treelet_bvh_mapping[addr + offset] in every record when the layout is remapped (:1682,:1763,...)
          const uint64_t address = SIMPLE ? simple_base + (uint64_t)slot * 64u
                                          : (p.remap ? __ldg(p.remap + slot) : (one_span ? span_host + (uint64_t)slot * 64u : slot_to_host(av, slot)) + (uint64_t)delta);
          const uint32_t type = code_type(code), size = code_size(code);
          const unsigned long long j = j0 + pos[u];     // < offsets[n_rays] <= out_capacity (checked on entry)
          if (PACKED) p.packed[j] = rec[u];
          else {
            *reinterpret_cast<uint4*>(p.txns + j) = make_uint4((uint32_t)address, (uint32_t)(address >> 32), size, type);
            p.tids[j] = tid[u];
          }
          // g_rt_mem_access_type[type]++ as eight 8-bit lanes in two words (types 0..3 | 4..7)
          const uint32_t inc = 1u << (8u * (type & 3u));
          packed_lo += (type & 4u) ? 0u : inc; packed_hi += (type & 4u) ? inc : 0u;
        }
      }
      since_flush += K3_ILP;
      if (since_flush > 255u - K3_ILP) {
#pragma unroll
        for (int c = 0; c < 4; c++) { hc[c] += (packed_lo >> (8 * c)) & 0xffu; hc[4 + c] += (packed_hi >> (8 * c)) & 0xffu; }
        packed_lo = packed_hi = 0; since_flush = 0;
      }
      if (p.treelet_hist && p.count) {
        // Treelet visit histogram.  The hot bins (the treelets at the top of the tree) receive a record from every
        // ray, and same-address atomics serialise in L2, so: (1) lanes hold consecutive records, runs of one treelet
        // are folded with a shuffle + ballot and only the run head adds; (2) the CTA accumulates into a small
        // shared-memory hash table and flushes it once at the end; only table collisions go straight to L2.
#pragma unroll
        for (int u = 0; u < K3_ILP; u++) {
          const bool a = valid[u] && tid[u] != VSRT_NO_TID;
          const uint32_t prev = __shfl_up_sync(full, tid[u], 1);
          const unsigned act = __ballot_sync(full, a);
          const bool head = a && (lane == 0 || !((act >> (lane - 1)) & 1u) || prev != tid[u]);
          const unsigned heads = __ballot_sync(full, head);
          if (head) {
            // run = lanes up to the next head or the first inactive lane
            const unsigned above = (lane == 31) ? 0u : ((heads | ~act) & (0xffffffffu << (lane + 1)));
            const uint32_t run = (above ? (uint32_t)(__ffs(above) - 1) : 32u) - (uint32_t)lane;
            const uint32_t h = (tid[u] * 2654435761u) >> (32 - TBITS);
            uint32_t old = t_key[h];                                             // hot treelets own their slot already: no CAS
            if (old == VSRT_NO_TID) old = atomicCAS(&t_key[h], VSRT_NO_TID, tid[u]);
            if (old == VSRT_NO_TID || old == tid[u]) atomicAdd(&t_cnt[h], run);
            else atomicAdd(p.treelet_hist + tid[u], (unsigned long long)run);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) { hc[c] += (packed_lo >> (8 * c)) & 0xffu; hc[4 + c] += (packed_hi >> (8 * c)) & 0xffu; }
  }
  if (!p.count) return;          // records only (a later full expansion of a batch that was first delivered packed)
  if (p.treelet_hist) {
    if (VSRT_K3_WARP_TABLE) {
      __syncwarp();
      for (uint32_t i = threadIdx.x & 31; i < (1u << TBITS); i += 32)
        if (t_key[i] != VSRT_NO_TID && t_cnt[i]) atomicAdd(p.treelet_hist + t_key[i], (unsigned long long)t_cnt[i]);
    } else {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < (1u << K3_HASH_BITS); i += K3_THREADS)
        if (s_hkey[i] != VSRT_NO_TID && s_hcnt[i]) atomicAdd(p.treelet_hist + s_hkey[i], (unsigned long long)s_hcnt[i]);
    }
  }
  uint32_t mine = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint32_t s = __reduce_add_sync(0xffffffffu, hc[c]);
    if (VSRT_K3_WARP_TABLE) { if ((threadIdx.x & 31) == c) mine = s; }
    else if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_hist[c], s);
  }
  if (VSRT_K3_WARP_TABLE) {   // lane c of every warp adds type c straight to the global counters
    const uint32_t c = threadIdx.x & 31;
    if (c < 8 && mine) {
      atomicAdd(p.counters->v + CI_TYPE0 + c, (unsigned long long)mine);
      atomicAdd(p.counters->v + CI_ACCESSED, (unsigned long long)mine * (c == VSRT_TXN_BVH_INSTANCE_LEAF ? 128ull : (c == VSRT_TXN_BVH_PRIMITIVE_LEAF_DESCRIPTOR ? 8ull : 64ull)));
    }
    return;
  }
  __syncthreads();
  if (threadIdx.x < 8 && s_hist[threadIdx.x]) {
    const unsigned long long n = s_hist[threadIdx.x];
    atomicAdd(p.counters->v + CI_TYPE0 + threadIdx.x, n);
    const unsigned long long bytes = n * (threadIdx.x == VSRT_TXN_BVH_INSTANCE_LEAF ? 128ull : (threadIdx.x == VSRT_TXN_BVH_PRIMITIVE_LEAF_DESCRIPTOR ? 8ull : 64ull));
    atomicAdd(p.counters->v + CI_ACCESSED, bytes);
  }
}

// ---------------------------------------------------------------- K3, ray-chunk formulation (variant: VSRT_K3_RAYS=1 in the environment)
// The window kernel above spends a fifth of its instructions finding out which ray a record belongs to and a third on the hash
// table of the treelet histogram.  Here a warp still owns 32 consecutive rays and their contiguous output range, but walks it as
// CHUNKS: chunk = 32 consecutive records of ONE ray (the last chunk of a ray is partly empty).  Ray and record index of a lane are
// then a popcount and an add -- no search, no bitmap -- and the staged read, the 16-byte store and the 4-byte store of a chunk
// stay single coalesced transactions.  The chunks of the warp's 32 rays are numbered through (a warp scan of chunks per ray), so
// VSRT_K3V2_ILP of them are in flight whatever the ray lengths.  The grid is persistent (CTAs stride over the ray blocks): the
// CTA-private histogram table is loaded and flushed once per CTA, not once per 128 rays, and it is direct-mapped with STATIC keys
// chosen by K0 (vsrt_internal.h, VSRT_HOT_N): a run of records of treelet t adds to the shared-memory counter of bucket
// t & (N - 1) if K0 gave that bucket to t, and to the global histogram otherwise -- one compare instead of a CAS protocol, and the
// treelets near the top of the tree, which every ray visits, can never lose their bucket to a treelet that happened to come first.
// MEASURED (B200, bench workload, profiles/README.md): 0.77 ms against 0.59 ms for the window kernel; without the histogram 0.60
// against 0.46 ms.  A chunk is two thirds full on average (42 records per ray), so the kernel issues 1.5 x the loads and stores
// for the same records, which costs more than the ray search it saves; and the deep treelets -- most records -- are shared by
// NEIGHBOURING rays, which a table private to 128 consecutive rays captures and a static top-of-tree table sends to L2 atomics.
// ILP 8: 0.86 ms, 10 CTAs per SM: 0.85 ms.  Kept as a variant; the window kernel stays the default.
#ifndef VSRT_K3_V2
#define VSRT_K3_V2 1
#endif
#ifndef VSRT_K3V2_MIN_BLOCKS
#define VSRT_K3V2_MIN_BLOCKS 8
#endif
#ifndef VSRT_K3V2_ILP
#define VSRT_K3V2_ILP 4
#endif
template <bool SIMPLE, bool PACKED>
__global__ void __launch_bounds__(K3_THREADS, VSRT_K3V2_MIN_BLOCKS) k_compact_rays(const CompactParams p) {
  if (p.err_flags && (*reinterpret_cast<const volatile uint32_t*>(p.err_flags) & p.fatal_mask)) return;
  if (p.offsets[p.n_rays] > p.out_capacity) return;
  constexpr int ILP = VSRT_K3V2_ILP;
  __shared__ unsigned int s_key[VSRT_HOT_N], s_cnt[VSRT_HOT_N];
  __shared__ unsigned int s_hist[8];
  const bool do_hist = p.treelet_hist != nullptr && p.count != 0;
  for (uint32_t i = threadIdx.x; i < VSRT_HOT_N; i += K3_THREADS) { s_key[i] = do_hist ? __ldg(p.tv.hot_keys + i) : VSRT_NO_TID; s_cnt[i] = 0; }
  if (threadIdx.x < 8) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const ArenaView& av = p.av;
  const bool one_span = av.n_spans == 1;
  const uint64_t span_host = one_span ? av.spans[0].host : 0ull;
  const uint64_t simple_base = span_host + (uint64_t)av.tlas_delta;
  uint32_t hc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };      // records per CODE seen by this lane (code 7 = TLAS internal is folded into type 1 at the end)
  unsigned long long pk = 0; uint32_t since_flush = 0;   // the same as eight 8-bit counters, spilled into hc[] before they can wrap
  const uint64_t n_blocks = (p.n_rays + K3_RAYS - 1) / K3_RAYS;
  for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const uint64_t rw0 = blk * K3_RAYS + (uint64_t)(threadIdx.x >> 5) * 32u;
    if (rw0 >= p.n_rays) continue;
    const uint32_t nr = (uint32_t)min((uint64_t)32, p.n_rays - rw0);
    const unsigned long long my_off = p.offsets[rw0 + min((uint32_t)lane, nr)];
    const unsigned long long j0 = __shfl_sync(full, my_off, 0);
    const uint32_t total = (uint32_t)(__shfl_sync(full, my_off, 31) - j0) + 0u;   // lane 31 holds offsets[rw0 + min(31, nr)]
    const uint32_t end = nr < 32u ? total : (uint32_t)(p.offsets[rw0 + 32u] - j0);  // records of the warp's rays
    const uint32_t rel = (uint32_t)lane < nr ? (uint32_t)(my_off - j0) : end;
    uint32_t nxt = __shfl_down_sync(full, rel, 1); if (lane == 31) nxt = end;
    const uint32_t cntl = nxt - rel;                                 // records of this lane's ray (0 beyond the last ray)
    const uint32_t nch = (cntl + 31u) >> 5;                          // its chunks
    uint32_t incl = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
    const uint32_t excl = incl - nch, n_chunks = __shfl_sync(full, incl, 31);
    const uint32_t* stage_w = p.stage + rw0 * (uint64_t)p.cap;
    for (uint32_t t0 = 0; t0 < n_chunks; t0 += ILP) {
      uint32_t ray[ILP], k[ILP], pos[ILP], rec[ILP], tid[ILP]; bool valid[ILP];
#pragma unroll
      for (int u = 0; u < ILP; u++) {
        const uint32_t t = t0 + (uint32_t)u;
        ray[u] = min((uint32_t)__popc(__ballot_sync(full, incl <= t)), 31u);     // the ray chunk t belongs to: all rays before it are complete
        const uint32_t ex = __shfl_sync(full, excl, (int)ray[u]), c = __shfl_sync(full, cntl, (int)ray[u]), o = __shfl_sync(full, rel, (int)ray[u]);
        k[u] = ((t - ex) << 5) + (uint32_t)lane; valid[u] = t < n_chunks && k[u] < c; pos[u] = o + k[u];
      }
#pragma unroll
      for (int u = 0; u < ILP; u++) rec[u] = valid[u] ? __ldg(stage_w + (ray[u] * p.cap + k[u])) : 0u;
#pragma unroll
      for (int u = 0; u < ILP; u++) { tid[u] = valid[u] ? __ldg(p.tv.node_tid + (rec[u] >> 3)) : VSRT_NO_TID; if (tid[u] != VSRT_NO_TID) tid[u] &= VSRT_TID_MASK; }
#pragma unroll
      for (int u = 0; u < ILP; u++) {
        if (valid[u]) {
          const uint32_t slot = rec[u] >> 3, code = rec[u] & 7u;
          const unsigned long long j = j0 + pos[u];     // < offsets[n_rays] <= out_capacity (checked on entry)
          if (PACKED) p.packed[j] = rec[u];
          else {
            int64_t delta = av.tlas_delta;                // host -> simulated-device offset the reference applies to this record
            if (!SIMPLE && !av.uniform_delta) {
              const uint32_t* seg = stage_w + (uint64_t)ray[u] * p.cap;
              if (code == C_STRUCT && k[u] > 0) { int64_t d; if (blas_delta_of(av, slot, d)) delta = d; }          // :1908-1913 / :2640-2645
              else if (p.mode == VSRT_MODE_DFS && code != C_INTERNAL_TLAS && code != C_INSTANCE && k[u] > 0) {
                // traceRay keeps device_offset = offset of the BLAS it is inside (:2640) until the next TLAS node (:2503,:2605)
                for (uint32_t b = k[u]; b-- > 0;) { const uint32_t pr = __ldg(seg + b); if ((pr & 7u) == C_STRUCT && b > 0) { int64_t d; if (blas_delta_of(av, pr >> 3, d)) delta = d; break; } }
              }
            }
            const uint64_t address = SIMPLE ? simple_base + (uint64_t)slot * 64u
                                            : (p.remap ? __ldg(p.remap + slot) : (one_span ? span_host + (uint64_t)slot * 64u : slot_to_host(av, slot)) + (uint64_t)delta);
            *reinterpret_cast<uint4*>(p.txns + j) = make_uint4((uint32_t)address, (uint32_t)(address >> 32), code_size(code), code_type(code));
            p.tids[j] = tid[u];
          }
          pk += 1ull << (code << 3);
        }
      }
      since_flush += ILP;
      if (since_flush > 255u - ILP) {
#pragma unroll
        for (int c = 0; c < 8; c++) hc[c] += (uint32_t)(pk >> (8 * c)) & 0xffu;
        pk = 0; since_flush = 0;
      }
      if (do_hist) {
#pragma unroll
        for (int u = 0; u < ILP; u++) {
          // lanes hold consecutive records of one ray: a run of one treelet is folded into its first lane
          const uint32_t prev = __shfl_up_sync(full, tid[u], 1);
          const bool first = lane == 0 || prev != tid[u];
          const unsigned bnd = __ballot_sync(full, first);
          if (first && tid[u] != VSRT_NO_TID) {
            const unsigned above = (bnd >> lane) >> 1;
            const uint32_t run = above ? (uint32_t)__ffs(above) : 32u - (uint32_t)lane;
            const uint32_t h = tid[u] & (VSRT_HOT_N - 1u);
            if (s_key[h] == tid[u]) atomicAdd(&s_cnt[h], run);
            else atomicAdd(p.treelet_hist + tid[u], (unsigned long long)run);
          }
        }
      }
    }
  }
  if (!p.count) return;          // records only (a later full expansion of a batch that was first delivered packed)
#pragma unroll
  for (int c = 0; c < 8; c++) hc[c] += (uint32_t)(pk >> (8 * c)) & 0xffu;
  hc[VSRT_TXN_BVH_INTERNAL_NODE] += hc[C_INTERNAL_TLAS]; hc[C_INTERNAL_TLAS] = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint32_t s = __reduce_add_sync(0xffffffffu, hc[c]);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_hist[c], s);
  }
  __syncthreads();
  if (do_hist)
    for (uint32_t i = threadIdx.x; i < VSRT_HOT_N; i += K3_THREADS)
      if (s_cnt[i]) atomicAdd(p.treelet_hist + s_key[i], (unsigned long long)s_cnt[i]);
  if (threadIdx.x < 8 && s_hist[threadIdx.x]) {
    const unsigned long long n = s_hist[threadIdx.x];
    atomicAdd(p.counters->v + CI_TYPE0 + threadIdx.x, n);
    const unsigned long long bytes = n * (threadIdx.x == VSRT_TXN_BVH_INSTANCE_LEAF ? 128ull : (threadIdx.x == VSRT_TXN_BVH_PRIMITIVE_LEAF_DESCRIPTOR ? 8ull : 64ull));
    atomicAdd(p.counters->v + CI_ACCESSED, bytes);
  }
}

__global__ void k_tid_to_addr(const ArenaView av, const TreeletView tv, const uint32_t* __restrict__ tids, uint64_t n, unsigned long long* __restrict__ out,
                              unsigned long long remap_base, unsigned long long remap_pitch) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = tids[i];
  if (t == VSRT_NO_TID || t >= tv.n_treelets) out[i] = ~0ull;
  else out[i] = remap_pitch ? remap_base + (unsigned long long)t * remap_pitch : slot_to_host(av, __ldg(tv.tl_root + t)) + (uint64_t)av.tlas_delta;
}

// Packed form of the trace for host consumers that expand records at the point of use: the staged words (slot << 3 | code)
// of every ray, CSR-compacted.  One warp per ray, lanes over its records (both sides coalesced).
__global__ void __launch_bounds__(256) k_pack_trace(const uint32_t* __restrict__ stage, uint32_t cap, const unsigned long long* __restrict__ offsets,
                                                    uint64_t n_rays, uint32_t* __restrict__ out, uint64_t out_capacity) {
  const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const unsigned long long o0 = offsets[r], o1 = offsets[r + 1];
  const uint32_t* seg = stage + r * (uint64_t)cap;
  for (unsigned long long k = threadIdx.x & 31u; o0 + k < o1 && o0 + k < out_capacity; k += 32u) out[o0 + k] = __ldg(seg + k);
}

// Node-visit histogram (optional): records per node address, per 64-byte slot.  One warp takes 32 consecutive rays and
// walks their staged records row by row (lane = ray, row = k-th record): near the top of the tree the lanes of a row name the
// same few nodes, so the row is grouped with __match_any_sync and only the first lane of each group adds, with the group's size
// -- the root gets one atomic per warp instead of one per ray.
__global__ void __launch_bounds__(256) k_node_hist(const uint32_t* __restrict__ stage, uint32_t cap, const unsigned long long* __restrict__ offsets, uint64_t n_rays,
                                                   unsigned long long* __restrict__ hist, const uint32_t* __restrict__ err_flags, uint32_t fatal_mask) {
  if (err_flags && (*reinterpret_cast<const volatile uint32_t*>(err_flags) & fatal_mask)) return;
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t cnt = r < n_rays ? (uint32_t)(offsets[r + 1] - offsets[r]) : 0u;
  const uint32_t rows = __reduce_max_sync(0xffffffffu, cnt);
  const uint32_t* seg = stage + r * (uint64_t)cap;
  for (uint32_t k = 0; k < rows; k++) {
    const bool v = k < cnt;
    const uint32_t slot = v ? (__ldg(seg + k) >> 3) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    if (v && lane == __ffs(peers) - 1) atomicAdd(hist + slot, (unsigned long long)__popc(peers));
  }
}

}  // namespace

int vsrt_launch_node_hist(const uint32_t* stage, uint32_t cap, const uint64_t* offsets, uint64_t n_rays, unsigned long long* hist, const uint32_t* err_flags,
                          uint32_t fatal_mask, cudaStream_t st) {
  if (n_rays == 0) return VSRT_OK;
  k_node_hist<<<(unsigned)((n_rays + 255) / 256), 256, 0, st>>>(stage, cap, (const unsigned long long*)offsets, n_rays, hist, err_flags, fatal_mask);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

size_t vsrt_scan_tmp_bytes(uint64_t n) { return ((n + SCAN_TILE - 1) / SCAN_TILE + 2) * sizeof(unsigned long long); }

int vsrt_launch_scan(const uint32_t* counts, uint64_t n, uint64_t* offsets, void* tmp, cudaStream_t st, unsigned long long* total_out) {
  const uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n_tiles == 0) { cudaMemsetAsync(offsets, 0, 8, st); if (total_out) cudaMemsetAsync(total_out, 0, 8, st); return VSRT_OK; }
  if (cudaMemsetAsync(tmp, 0, (size_t)(n_tiles + 1) * 8, st) != cudaSuccess) return VSRT_E_CUDA;
  k_scan_onepass<<<(unsigned)n_tiles, SCAN_THREADS, 0, st>>>(counts, n, (unsigned long long*)tmp, (unsigned long long*)offsets, total_out);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

template <bool SIMPLE, bool PACKED>
static int launch_compact_rays(const CompactParams& p, uint64_t n_blocks, cudaStream_t st) {
  static int s_blocks[64], s_sm[64];      // occupancy of this instantiation, per device
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (s_blocks[dev] == 0) {
    cudaDeviceGetAttribute(&s_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_blocks[dev], k_compact_rays<SIMPLE, PACKED>, K3_THREADS, 0) != cudaSuccess || s_blocks[dev] < 1) s_blocks[dev] = 4;
  }
  const unsigned grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)s_blocks[dev] * (uint64_t)s_sm[dev]);
  k_compact_rays<SIMPLE, PACKED><<<grid, K3_THREADS, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_compact(const CompactParams& p, cudaStream_t st) {
  const uint64_t n_blocks = (p.n_rays + K3_RAYS - 1) / K3_RAYS;
  if (n_blocks == 0) return VSRT_OK;
  // the window kernel is the default; VSRT_K3_RAYS=1 picks the ray-chunk kernel (measured: 0.77 vs 0.59 ms, see its header)
  static const bool chunks = VSRT_K3_V2 && getenv("VSRT_K3_RAYS") && atoi(getenv("VSRT_K3_RAYS")) != 0;
  if (chunks) {
    if (p.packed) return launch_compact_rays<true, true>(p, n_blocks, st);
    if (p.av.n_spans == 1 && p.av.uniform_delta && !p.remap) return launch_compact_rays<true, false>(p, n_blocks, st);
    return launch_compact_rays<false, false>(p, n_blocks, st);
  }
  int n_sm = 148;
  if (VSRT_K3_PERSIST) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  int per_sm = K3_CTAS_PER_SM;
  if (VSRT_K3_PERSIST && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_compact<false, false>, K3_THREADS, 0) != cudaSuccess || per_sm < 1)) per_sm = 1;
  const uint64_t grid = VSRT_K3_PERSIST ? std::min<uint64_t>(n_blocks, (uint64_t)n_sm * (uint64_t)per_sm) : n_blocks;
  if (p.packed) k_compact<true, true><<<(unsigned)grid, K3_THREADS, 0, st>>>(p);      // addresses are not formed at all
  else if (p.av.n_spans == 1 && p.av.uniform_delta && !p.remap) k_compact<true, false><<<(unsigned)grid, K3_THREADS, 0, st>>>(p);
  else k_compact<false, false><<<(unsigned)grid, K3_THREADS, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_tid_to_addr(const ArenaView& av, const TreeletView& tv, const uint32_t* tids, uint64_t n, uint64_t* out,
                            uint64_t remap_base, uint64_t remap_pitch, cudaStream_t st) {
  if (n == 0) return VSRT_OK;
  k_tid_to_addr<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(av, tv, tids, n, (unsigned long long*)out, remap_base, remap_pitch);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

int vsrt_launch_pack_trace(const uint32_t* stage, uint32_t cap, const uint64_t* offsets, uint64_t n_rays, uint32_t* out, uint64_t out_capacity, cudaStream_t st) {
  if (n_rays == 0) return VSRT_OK;
  k_pack_trace<<<(unsigned)((n_rays + 7) / 8), 256, 0, st>>>(stage, cap, (const unsigned long long*)offsets, n_rays, out, out_capacity);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}
