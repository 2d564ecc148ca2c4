// rayorder.cu -- the order in which K1's lanes pick up the rays of a batch.
//
// The reference traces the lanes of a warp one after the other (abstract_hardware_model.cc:3052-3063), so nothing about a ray's
// results depends on which rays run beside it: K1 may hand its lanes the rays in any order as long as every output stays indexed
// by the ray's own id (staging segment, hit record, count -- all are; K3 never sees the order).  For an incoherent batch
// (diffuse bounces, BASELINE.json configs[2..4]) the input order puts 32 unrelated rays into a warp: they are at different nodes
// of different subtrees, every node fetch is a separate line and the warp serialises over its phases.  Sorting the batch by a
// space-filling key -- Morton code of the origin cell (7 bits per axis over the batch's own origin box) followed by the direction
// octant -- gives a warp neighbouring rays that walk the same part of the tree.  Origin first: rays that start close together
// share the whole chain of boxes around their origin whatever their directions; measured on bounce rays, a key with coarser
// origin cells than the input's own neighbourhoods makes K1 slower, which is why AUTO mode (vsrt_capi.cu) sorts only batches
// whose consecutive rays are far apart.
//
//   k_ray_bounds   origin box of the batch
//   k_ray_keys     24-bit key per ray
//   3 x (count, scan, scatter)   stable LSD radix sort of (key, ray id), 8 bits per pass, 2048 keys per CTA
//
// Every kernel after k_ray_bounds returns at once when the device-side decision word says "keep the input order" (a batch whose
// origins all coincide has no origin box to sort by, unless `force`); K1 reads the same word.
#include "vsrt_device.cuh"
#include <algorithm>

namespace {

constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32, RS_BINS = 256;

__device__ __forceinline__ unsigned int f2ord(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }   // order-preserving
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

// ext[0..2] min origin, [3..5] max origin (ordered-uint encoding); ext[6] = decision (1 sort, 0 keep) written by the last block
__global__ void __launch_bounds__(256) k_ray_bounds(const vsrt_ray* __restrict__ rays, uint64_t n, unsigned int* __restrict__ ext, unsigned int* __restrict__ done,
                                                    uint32_t force) {
  unsigned int lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0, 0, 0 };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) { const float v = __ldg(&rays[i].origin[a]); if (v == v) { const unsigned int o = f2ord(v); lo[a] = min(lo[a], o); hi[a] = max(hi[a], o); } }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) { lo[a] = __reduce_min_sync(0xffffffffu, lo[a]); hi[a] = __reduce_max_sync(0xffffffffu, hi[a]); }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(ext + a, lo[a]); atomicMax(ext + 3 + a, hi[a]); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      __threadfence();
      // one origin for the whole batch = a camera batch in scan order: its neighbours in the input are its neighbours in space
      const volatile unsigned int* e = ext;
      const bool spread = e[0] < e[3] || e[1] < e[4] || e[2] < e[5];
      ext[6] = (force || spread) ? 1u : 0u;
    }
  }
}

__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // 8 bits -> every third bit
  v &= 0xffu; v = (v | (v << 8)) & 0x00F00Fu; v = (v | (v << 4)) & 0x0C30C3u; v = (v | (v << 2)) & 0x249249u; return v;
}

__global__ void __launch_bounds__(256) k_ray_keys(const vsrt_ray* __restrict__ rays, uint64_t n, const unsigned int* __restrict__ ext,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
  if (ext[6] == 0u) return;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t q[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float lo = ord2f(ext[a]), hi = ord2f(ext[3 + a]);
    const float v = __ldg(&rays[i].origin[a]);
    const float w = hi > lo ? (v - lo) / (hi - lo) : 0.0f;            // ordering only: no bit-exactness contract here
    q[a] = (uint32_t)fminf(fmaxf(w * 128.0f, 0.0f), 127.0f);
    d[a] = __float_as_uint(__ldg(&rays[i].direction[a])) >> 31;
  }
  const uint32_t om = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);   // 21 bits
  keys[i] = (om << 3) | d[0] | (d[1] << 1) | (d[2] << 2);
  ids[i] = (uint32_t)i;
}

// digit-major table: cnt[digit * n_blocks + block]
__global__ void __launch_bounds__(RS_THREADS) k_radix_count(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t n_blocks, uint32_t* __restrict__ cnt,
                                                            const unsigned int* __restrict__ gate) {
  if (gate && *gate == 0u) return;
  __shared__ unsigned int h[RS_BINS];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) { const uint64_t k = base + (uint64_t)i * RS_THREADS + threadIdx.x; if (k < n) atomicAdd(&h[(keys[k] >> shift) & 255u], 1u); }
  __syncthreads();
  cnt[(uint64_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter.  Warp w owns keys [base + 256 w, base + 256 (w + 1)) of the tile, lane l item i is element 32 i + l of that
// segment, so "order" = (warp, item, lane).  Rank of an element among the equal digits before it: the warp's running count of
// the digit (shared memory, bumped once per round by the lowest lane of each match group) + its position inside the group.
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ids, uint64_t n, int shift, uint32_t n_blocks,
                                                              const unsigned long long* __restrict__ off, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ ids_out,
                                                              const unsigned int* __restrict__ gate) {
  if (gate && *gate == 0u) return;
  __shared__ unsigned int wc[RS_WARPS][RS_BINS];
  for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wc[0][0])[i] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t seg = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)w * (32 * RS_ITEMS);
  uint32_t key[RS_ITEMS], id[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const uint64_t k = seg + (uint64_t)i * 32 + lane;
    const bool v = k < n;
    key[i] = v ? keys[k] : 0xFFFFFFFFu; id[i] = v ? ids[k] : 0u;
    const uint32_t dgt = (key[i] >> shift) & 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, v ? dgt : 0x100u);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (v && before == 0) { old = wc[w][dgt]; wc[w][dgt] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
    rank[i] = old + before;
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over the warps, per digit (thread d handles digit d), plus the tile's global offset for the digit
  {
    const unsigned long long g = off[(uint64_t)threadIdx.x * n_blocks + blockIdx.x];
    unsigned int run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) { const unsigned int c = wc[ww][threadIdx.x]; wc[ww][threadIdx.x] = (unsigned int)g + run; run += c; }   // n < 2^32
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const uint64_t k = seg + (uint64_t)i * 32 + lane;
    if (k < n) { const uint32_t dgt = (key[i] >> shift) & 255u; const uint32_t pos = wc[w][dgt] + rank[i]; keys_out[pos] = key[i]; ids_out[pos] = id[i]; }
  }
}

}  // namespace

// ---- stable LSD radix sort of (key, id) pairs, 8 bits per pass, shared with the treelet-binned traversal (traverse_tb.cu)
size_t vsrt_radix_tmp_bytes(uint64_t n) {
  const uint64_t n_blocks = (n + RS_TILE - 1) / RS_TILE, m = n_blocks * RS_BINS;
  return ((m + 63) & ~63ull) * 4 + (m + 1) * 8 + 256 + vsrt_scan_tmp_bytes(m) + 256;
}
// Sorts by the low 8 * passes bits.  (kA, iA) hold the input, (kB, iB) are scratch of the same size; *keys_out / *ids_out name
// the buffers the result ends up in.  `gate` (device word, may be NULL): every kernel returns at once when it is zero.
int vsrt_launch_radix_sort(uint32_t* kA, uint32_t* iA, uint32_t* kB, uint32_t* iB, uint64_t n, int passes, void* tmp, const unsigned int* gate,
                           uint32_t** keys_out, uint32_t** ids_out, cudaStream_t st) {
  const uint64_t n_blocks = (n + RS_TILE - 1) / RS_TILE, m = n_blocks * RS_BINS;
  uint8_t* p = (uint8_t*)tmp;
  uint32_t* cnt = (uint32_t*)p; p += ((m + 63) & ~63ull) * 4;
  unsigned long long* off = (unsigned long long*)p; p += (m + 1) * 8;
  p = (uint8_t*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  void* scan_tmp = p;
  uint32_t* ks = kA; uint32_t* is = iA; uint32_t* kd = kB; uint32_t* id = iB;
  for (int pass = 0; pass < passes && n; pass++) {
    k_radix_count<<<(unsigned)n_blocks, RS_THREADS, 0, st>>>(ks, n, 8 * pass, (uint32_t)n_blocks, cnt, gate);
    const int rc = vsrt_launch_scan(cnt, m, (uint64_t*)off, scan_tmp, st); if (rc) return rc;
    k_radix_scatter<<<(unsigned)n_blocks, RS_THREADS, 0, st>>>(ks, is, n, 8 * pass, (uint32_t)n_blocks, off, kd, id, gate);
    std::swap(ks, kd); std::swap(is, id);
  }
  if (keys_out) *keys_out = ks;
  if (ids_out) *ids_out = is;
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

size_t vsrt_rayorder_tmp_bytes(uint64_t n) {
  // ext (8 words) | keys A, ids A, keys B, ids B | radix scratch
  return 64 + 4 * ((n + 63) & ~63ull) * 4 + vsrt_radix_tmp_bytes(n);
}

// Leaves the sorted ray ids in *perm_out (inside tmp) and the decision word in *decision_out; both device pointers.
int vsrt_launch_rayorder(const vsrt_ray* rays_dev, uint64_t n, bool force, void* tmp, const uint32_t** perm_out, const uint32_t** decision_out, cudaStream_t st) {
  const uint64_t np = (n + 63) & ~63ull;
  uint8_t* p = (uint8_t*)tmp;
  unsigned int* ext = (unsigned int*)p; p += 64;
  uint32_t* kA = (uint32_t*)p; p += np * 4; uint32_t* iA = (uint32_t*)p; p += np * 4;
  uint32_t* kB = (uint32_t*)p; p += np * 4; uint32_t* iB = (uint32_t*)p; p += np * 4;
  const unsigned int init[8] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u };
  if (cudaMemcpyAsync(ext, init, sizeof(init), cudaMemcpyHostToDevice, st) != cudaSuccess) return VSRT_E_CUDA;
  const unsigned bgrid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
  k_ray_bounds<<<bgrid, 256, 0, st>>>(rays_dev, n, ext, ext + 7, force ? 1u : 0u);
  k_ray_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rays_dev, n, ext, kA, iA);
  uint32_t* is = nullptr;
  const int rc = vsrt_launch_radix_sort(kA, iA, kB, iB, n, 3, p, ext + 6, nullptr, &is, st);
  if (rc) return rc;
  *perm_out = is; *decision_out = ext + 6;
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}
