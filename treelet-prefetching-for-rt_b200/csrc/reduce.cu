// reduce.cu -- the path's only multi-GPU exchange: rays are sharded over ranks, one context per GPU with the BVH
// and the treelet tables replicated, and what has to be combined are the functional counters (cuda-sim.h:155-166,
// vulkan_ray_tracing.cc:1658,1687,2214,2260,2269-2282) and the visit histograms the treelet prefetcher's popularity vote reads
// (gpgpu-sim/shader.cc:3424-3433).  Everything here sits behind the C-ABI (vsrt_comm_* / vsrt_reduce_counters), so a C++ host
// -- the reference's simulator -- can run N contexts without any Python:
//
//   snapshot (traversal stream)   what this rank added since the previous reduce: 16 SUM counters as deltas, the 2 MAX counters
//                                 in this rank's own pair of an n_ranks-wide table (zeros elsewhere, so that a SUM reduce
//                                 delivers every rank's value), the per-treelet / per-node visit deltas as u32
//   reduce   (reduce stream)      ONE ncclGroup: ncclAllReduce(SUM, u64) over the header and ncclAllReduce(SUM, u32) over the
//                                 histogram deltas, in place in buffers the library owns (two sets, used alternately)
//   fold     (reduce stream)      global totals += reduced deltas; MAX over the per-rank pairs; a reduced record count of 2^32
//                                 or more means a u32 bin may have wrapped and is reported by vsrt_reduced_get
//
// The reduce of frame i runs beside the traversal of frame i + 1; a buffer set is reused only after the traversal stream has
// waited for the fold that last read it.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already loaded in the
// process, e.g. PyTorch's, else the system one), so libvsrt.so has no link-time dependency and single-GPU users never load it.
#include "vsrt_context.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  char err[256] = "";
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.lib) return true;
  static_assert(sizeof(ncclUniqueId) == VSRT_COMM_ID_BYTES, "VSRT_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  const char* names[] = { getenv("VSRT_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
  void* h = nullptr;
  for (const char* n : names) if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses
  for (const char* n : names) if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { snprintf(g_nccl.err, sizeof(g_nccl.err), "libnccl.so.2 not found (%s); set VSRT_NCCL_LIB", dlerror()); return false; }
#define SYM(field, name) do { *(void**)&g_nccl.field = dlsym(h, name); if (!g_nccl.field) { snprintf(g_nccl.err, sizeof(g_nccl.err), "%s missing from the NCCL library", name); return false; } } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy"); SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString"); SYM(GetVersion, "ncclGetVersion");
#undef SYM
  g_nccl.lib = h;
  return true;
}

// header of one reduce, u64 words: [0, 16) SUM deltas | [16, 16 + 2 * n_ranks) the MAX pairs | records delta | overflow marks
constexpr uint32_t H_SUM = VSRT_COUNTERS_N_SUM, H_MAX = VSRT_COUNTERS_N_MAX;
__host__ __device__ inline uint32_t hdr_words(uint32_t n_ranks) { return H_SUM + H_MAX * n_ranks + 2u; }

__global__ void __launch_bounds__(256) k_snapshot(const DevCounters* __restrict__ now, DevCounters* __restrict__ prev, unsigned long long* __restrict__ hdr,
                                                  uint32_t n_ranks, uint32_t rank, const unsigned long long* __restrict__ hist,
                                                  unsigned long long* __restrict__ hist_prev, uint32_t* __restrict__ dh, uint32_t n_hist, uint32_t with_header) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && with_header) {
    const uint32_t hw = hdr_words(n_ranks);
    for (uint32_t k = threadIdx.x; k + 1u < hw; k += blockDim.x) {   // the last word (overflow mark) was zeroed by the launcher: other blocks raise it
      unsigned long long v = 0;
      if (k < H_SUM) { v = now->v[k] - prev->v[k]; }
      else if (k < H_SUM + H_MAX * n_ranks) { const uint32_t r = (k - H_SUM) / H_MAX, f = (k - H_SUM) % H_MAX; v = r == rank ? now->v[H_SUM + f] : 0ull; }
      else if (k == hw - 2u) { for (int t = 0; t < 9; t++) v += now->v[CI_TYPE0 + t] - prev->v[CI_TYPE0 + t]; }   // records this rank adds
      hdr[k] = v;
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < H_SUM; k += blockDim.x) prev->v[k] = now->v[k];
  }
  if (i < n_hist) {
    const unsigned long long h = hist[i], d = h - hist_prev[i];
    hist_prev[i] = h;
    dh[i] = (uint32_t)d;
    if (d >> 32) atomicMax(hdr + hdr_words(n_ranks) - 1u, 1ull);    // does not fit 32 bits: reported by vsrt_reduced_get, never silent
  }
}

__global__ void __launch_bounds__(256) k_fold(const unsigned long long* __restrict__ hdr, uint32_t n_ranks, DevCounters* __restrict__ g, const uint32_t* __restrict__ dh,
                                              unsigned long long* __restrict__ g_hist, uint32_t n_hist, uint32_t* __restrict__ flags, uint32_t with_header) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (!with_header) { if (i < n_hist) g_hist[i] += dh[i]; return; }
  if (i < H_SUM) g->v[i] += hdr[i];
  else if (i < H_SUM + H_MAX) {
    unsigned long long m = g->v[i];
    for (uint32_t r = 0; r < n_ranks; r++) { const unsigned long long v = hdr[H_SUM + H_MAX * r + (i - H_SUM)]; m = v > m ? v : m; }
    g->v[i] = m;
  } else if (i == H_SUM + H_MAX) {
    const uint32_t hw = hdr_words(n_ranks);
    if ((hdr[hw - 2u] >> 32) || hdr[hw - 1u]) atomicOr(flags, 1u);   // 2^32 or more records between two reduces: a u32 bin may have wrapped
  }
  if (i < n_hist) g_hist[i] += dh[i];
}

}  // namespace

struct CommState {
  ncclComm_t comm = nullptr; bool owned = false;
  uint32_t n_ranks = 1, rank = 0;
  cudaStream_t rstream = nullptr;
  cudaEvent_t ready[2] = { nullptr, nullptr }, done[2] = { nullptr, nullptr };
  bool used[2] = { false, false };
  int turn = 0;
  DevBuf<unsigned long long> hdr[2]; DevBuf<uint32_t> dh[2];
  DevCounters* prev = nullptr; DevCounters* global = nullptr; uint32_t* flags = nullptr;
  DevBuf<unsigned long long> hist_prev, g_hist; uint32_t n_hist = 0;
  // the optional node-visit histogram (vsrt_enable_node_histogram): same scheme, per 64-byte slot
  DevBuf<uint32_t> ndh[2]; DevBuf<unsigned long long> node_prev, g_node; uint32_t n_node = 0;
  uint64_t n_reduces = 0;
};

namespace {

#define CUDA_OK(c, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return vsrt_fail(c, VSRT_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); } while (0)
#define NCCL_OK(c, x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return vsrt_fail(c, VSRT_E_COMM, "%s failed: %s", #x, g_nccl.GetErrorString(r_)); } while (0)

int comm_setup(vsrt_context* c, ncclComm_t comm, bool owned, uint32_t n_ranks, uint32_t rank) {
  CommState* s = new CommState();
  s->comm = comm; s->owned = owned; s->n_ranks = n_ranks; s->rank = rank;
  c->comm = s;
  CUDA_OK(c, cudaStreamCreateWithFlags(&s->rstream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    CUDA_OK(c, cudaEventCreateWithFlags(&s->ready[i], cudaEventDisableTiming)); CUDA_OK(c, cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming));
    CUDA_OK(c, s->hdr[i].ensure(hdr_words(n_ranks)));
  }
  CUDA_OK(c, cudaMalloc(&s->prev, sizeof(DevCounters))); CUDA_OK(c, cudaMalloc(&s->global, sizeof(DevCounters))); CUDA_OK(c, cudaMalloc(&s->flags, 4));
  // the reduce covers what is traced from now on
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  CUDA_OK(c, cudaMemcpy(s->prev, c->d_counters, sizeof(DevCounters), cudaMemcpyDeviceToDevice));
  CUDA_OK(c, cudaMemset(s->global, 0, sizeof(DevCounters))); CUDA_OK(c, cudaMemset(s->flags, 0, 4));
  return VSRT_OK;
}

// (re)size the histogram side for the current treelet tables.  Baseline: at vsrt_comm_init what the rank's histogram holds at that
// moment (the reduce covers what is traced from then on); when the tables appear or change later, zero -- formation zeroes the
// rank's histogram, and everything counted since belongs to the next reduce.
int comm_hist_setup(vsrt_context* c, bool at_init = false) {
  CommState* s = c->comm;
  const uint32_t n = c->hist_n;
  if (s->n_hist == n && s->g_hist.p) return VSRT_OK;
  CUDA_OK(c, cudaStreamSynchronize(s->rstream)); CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 2; i++) CUDA_OK(c, s->dh[i].ensure(std::max<uint32_t>(n, 1)));
  CUDA_OK(c, s->hist_prev.ensure(std::max<uint32_t>(n, 1))); CUDA_OK(c, s->g_hist.ensure(std::max<uint32_t>(n, 1)));
  if (n) {
    if (at_init) CUDA_OK(c, cudaMemcpy(s->hist_prev.p, c->d_hist.p, (size_t)n * 8, cudaMemcpyDeviceToDevice));
    else CUDA_OK(c, cudaMemset(s->hist_prev.p, 0, (size_t)n * 8));
    CUDA_OK(c, cudaMemset(s->g_hist.p, 0, (size_t)n * 8));
  }
  s->n_hist = n;
  return VSRT_OK;
}
int comm_node_setup(vsrt_context* c, bool at_init = false) {
  CommState* s = c->comm;
  const uint32_t n = c->node_hist_on ? c->node_hist_n : 0u;
  if (s->n_node == n) return VSRT_OK;
  CUDA_OK(c, cudaStreamSynchronize(s->rstream)); CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 2; i++) CUDA_OK(c, s->ndh[i].ensure(std::max<uint32_t>(n, 1)));
  CUDA_OK(c, s->node_prev.ensure(std::max<uint32_t>(n, 1))); CUDA_OK(c, s->g_node.ensure(std::max<uint32_t>(n, 1)));
  // (the node histogram is allocated, zeroed, by the first batch traced with it on: whatever it holds now belongs to this reduce)
  if (n) {
    if (at_init) CUDA_OK(c, cudaMemcpy(s->node_prev.p, c->d_node_hist.p, (size_t)n * 8, cudaMemcpyDeviceToDevice));
    else CUDA_OK(c, cudaMemset(s->node_prev.p, 0, (size_t)n * 8));
    CUDA_OK(c, cudaMemset(s->g_node.p, 0, (size_t)n * 8));
  }
  s->n_node = n;
  return VSRT_OK;
}

}  // namespace

void vsrt_comm_release(vsrt_context* c) {
  CommState* s = c->comm; if (!s) return;
  if (s->rstream) cudaStreamSynchronize(s->rstream);
  if (s->owned && s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  for (int i = 0; i < 2; i++) { if (s->ready[i]) cudaEventDestroy(s->ready[i]); if (s->done[i]) cudaEventDestroy(s->done[i]); s->hdr[i].release(); s->dh[i].release(); }
  s->hist_prev.release(); s->g_hist.release(); s->node_prev.release(); s->g_node.release(); s->ndh[0].release(); s->ndh[1].release();
  cudaFree(s->prev); cudaFree(s->global); cudaFree(s->flags);
  if (s->rstream) cudaStreamDestroy(s->rstream);
  delete s; c->comm = nullptr;
}

// vsrt_reset_counters: the rank's counters and histogram restart at zero, and so do the global totals
int vsrt_comm_counters_reset(vsrt_context* c) {
  CommState* s = c->comm; if (!s) return VSRT_OK;
  CUDA_OK(c, cudaStreamSynchronize(s->rstream));
  CUDA_OK(c, cudaMemset(s->prev, 0, sizeof(DevCounters))); CUDA_OK(c, cudaMemset(s->global, 0, sizeof(DevCounters))); CUDA_OK(c, cudaMemset(s->flags, 0, 4));
  if (s->n_hist != 0xFFFFFFFFu && s->n_hist) { CUDA_OK(c, cudaMemset(s->hist_prev.p, 0, (size_t)s->n_hist * 8)); CUDA_OK(c, cudaMemset(s->g_hist.p, 0, (size_t)s->n_hist * 8)); }
  if (s->n_node) { CUDA_OK(c, cudaMemset(s->node_prev.p, 0, (size_t)s->n_node * 8)); CUDA_OK(c, cudaMemset(s->g_node.p, 0, (size_t)s->n_node * 8)); }
  return VSRT_OK;
}

void vsrt_comm_treelets_changed(vsrt_context* c) { if (c->comm) c->comm->n_hist = 0xFFFFFFFFu; }   // forces comm_hist_setup at the next reduce

extern "C" {

int vsrt_comm_unique_id(uint8_t id[VSRT_COMM_ID_BYTES]) {
  if (!id) return VSRT_E_INVALID;
  if (!load_nccl()) return vsrt_fail(nullptr, VSRT_E_COMM, "%s", g_nccl.err);
  ncclUniqueId u;
  NCCL_OK(nullptr, g_nccl.GetUniqueId(&u));
  memcpy(id, &u, VSRT_COMM_ID_BYTES);
  return VSRT_OK;
}

int vsrt_comm_init(vsrt_context* c, uint32_t n_ranks, uint32_t rank, const uint8_t id[VSRT_COMM_ID_BYTES]) {
  if (!c || !id || n_ranks == 0 || rank >= n_ranks) return c ? vsrt_fail(c, VSRT_E_INVALID, "vsrt_comm_init: bad rank %u of %u", rank, n_ranks) : VSRT_E_INVALID;
  if (c->comm) return vsrt_fail(c, VSRT_E_INVALID, "this context already has a communicator");
  if (!load_nccl()) return vsrt_fail(c, VSRT_E_COMM, "%s", g_nccl.err);
  cudaSetDevice(c->device);
  ncclUniqueId u; memcpy(&u, id, VSRT_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  NCCL_OK(c, g_nccl.CommInitRank(&comm, (int)n_ranks, u, (int)rank));
  int rc = comm_setup(c, comm, true, n_ranks, rank); if (rc) return rc;
  rc = comm_hist_setup(c, true); if (rc) return rc;
  return comm_node_setup(c, true);
}

int vsrt_comm_attach(vsrt_context* c, void* nccl_comm, uint32_t n_ranks, uint32_t rank) {
  if (!c || !nccl_comm || n_ranks == 0 || rank >= n_ranks) return c ? vsrt_fail(c, VSRT_E_INVALID, "vsrt_comm_attach: bad arguments") : VSRT_E_INVALID;
  if (c->comm) return vsrt_fail(c, VSRT_E_INVALID, "this context already has a communicator");
  if (!load_nccl()) return vsrt_fail(c, VSRT_E_COMM, "%s", g_nccl.err);
  cudaSetDevice(c->device);
  int rc = comm_setup(c, (ncclComm_t)nccl_comm, false, n_ranks, rank); if (rc) return rc;
  rc = comm_hist_setup(c, true); if (rc) return rc;
  return comm_node_setup(c, true);
}

int vsrt_comm_destroy(vsrt_context* c) {
  if (!c) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  vsrt_comm_release(c);
  return VSRT_OK;
}

int vsrt_reduce_counters(vsrt_context* c, void* stream) {
  if (!c) return VSRT_E_INVALID;
  CommState* s = c->comm;
  if (!s) return vsrt_fail(c, VSRT_E_INVALID, "no communicator: call vsrt_comm_init / vsrt_comm_attach first");
  cudaSetDevice(c->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  int rc = comm_hist_setup(c); if (rc) return rc;
  rc = comm_node_setup(c); if (rc) return rc;
  const int b = s->turn; s->turn ^= 1;
  const uint32_t n = s->n_hist, nn = s->n_node, hw = hdr_words(s->n_ranks);
  // the buffer set may still be read by the fold of two reduces ago
  if (s->used[b]) CUDA_OK(c, cudaStreamWaitEvent(st, s->done[b], 0));
  CUDA_OK(c, cudaMemsetAsync(s->hdr[b].p + hw - 1u, 0, 8, st));
  const uint32_t grid = std::max<uint32_t>(1u, (n + 255u) / 256u);
  k_snapshot<<<grid, 256, 0, st>>>(c->d_counters, s->prev, s->hdr[b].p, s->n_ranks, s->rank, c->d_hist.p, s->hist_prev.p, s->dh[b].p, n, 1u);
  if (nn) k_snapshot<<<(nn + 255u) / 256u, 256, 0, st>>>(c->d_counters, s->prev, s->hdr[b].p, s->n_ranks, s->rank, c->d_node_hist.p, s->node_prev.p, s->ndh[b].p, nn, 0u);
  CUDA_OK(c, cudaGetLastError());
  CUDA_OK(c, cudaEventRecord(s->ready[b], st));
  CUDA_OK(c, cudaStreamWaitEvent(s->rstream, s->ready[b], 0));
  NCCL_OK(c, g_nccl.GroupStart());
  ncclResult_t r1 = g_nccl.AllReduce(s->hdr[b].p, s->hdr[b].p, hw, ncclUint64, ncclSum, s->comm, s->rstream);
  ncclResult_t r2 = n ? g_nccl.AllReduce(s->dh[b].p, s->dh[b].p, n, ncclUint32, ncclSum, s->comm, s->rstream) : ncclSuccess;
  ncclResult_t r2b = nn ? g_nccl.AllReduce(s->ndh[b].p, s->ndh[b].p, nn, ncclUint32, ncclSum, s->comm, s->rstream) : ncclSuccess;
  ncclResult_t r3 = g_nccl.GroupEnd();
  NCCL_OK(c, r1); NCCL_OK(c, r2); NCCL_OK(c, r2b); NCCL_OK(c, r3);
  const uint32_t fgrid = (std::max<uint32_t>(n, H_SUM + H_MAX + 1u) + 255u) / 256u;
  k_fold<<<fgrid, 256, 0, s->rstream>>>(s->hdr[b].p, s->n_ranks, s->global, s->dh[b].p, s->g_hist.p, n, s->flags, 1u);
  if (nn) k_fold<<<(nn + 255u) / 256u, 256, 0, s->rstream>>>(s->hdr[b].p, s->n_ranks, s->global, s->ndh[b].p, s->g_node.p, nn, s->flags, 0u);
  CUDA_OK(c, cudaGetLastError());
  CUDA_OK(c, cudaEventRecord(s->done[b], s->rstream));
  s->used[b] = true; s->n_reduces++;
  return VSRT_OK;
}

int vsrt_reduce_wait(vsrt_context* c, void* stream) {
  if (!c) return VSRT_E_INVALID;
  CommState* s = c->comm;
  if (!s) return vsrt_fail(c, VSRT_E_INVALID, "no communicator");
  cudaSetDevice(c->device);
  if (stream == (void*)(intptr_t)-1) { CUDA_OK(c, cudaStreamSynchronize(s->rstream)); return VSRT_OK; }
  cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
  for (int i = 0; i < 2; i++) if (s->used[i]) CUDA_OK(c, cudaStreamWaitEvent(st, s->done[i], 0));
  return VSRT_OK;
}

int vsrt_reduced_get(vsrt_context* c, vsrt_counters* out, uint64_t* treelet_hist, uint64_t capacity) {
  if (!c) return VSRT_E_INVALID;
  CommState* s = c->comm;
  if (!s) return vsrt_fail(c, VSRT_E_INVALID, "no communicator");
  cudaSetDevice(c->device);
  CUDA_OK(c, cudaStreamSynchronize(s->rstream));
  uint32_t flags = 0;
  CUDA_OK(c, cudaMemcpy(&flags, s->flags, 4, cudaMemcpyDeviceToHost));
  if (out) CUDA_OK(c, cudaMemcpy(out, s->global, sizeof(DevCounters), cudaMemcpyDeviceToHost));
  if (treelet_hist) {
    if (s->n_hist == 0xFFFFFFFFu || capacity < s->n_hist) return vsrt_fail(c, VSRT_E_CAPACITY, "the reduced histogram has %u entries", s->n_hist == 0xFFFFFFFFu ? 0u : s->n_hist);
    if (s->n_hist) CUDA_OK(c, cudaMemcpy(treelet_hist, s->g_hist.p, (size_t)s->n_hist * 8, cudaMemcpyDeviceToHost));
  }
  if (flags & 1u) return vsrt_fail(c, VSRT_E_CAPACITY, "2^32 or more records were traced between two reduces: the 32-bit histogram deltas may have wrapped; reduce more often");
  return VSRT_OK;
}

int vsrt_reduced_get_node_histogram(vsrt_context* c, uint64_t* visits_of_slot, uint64_t capacity) {
  if (!c || !visits_of_slot) return VSRT_E_INVALID;
  CommState* s = c->comm;
  if (!s) return vsrt_fail(c, VSRT_E_INVALID, "no communicator");
  if (!s->n_node) return vsrt_fail(c, VSRT_E_INVALID, "the node-visit histogram was not part of any reduce: vsrt_enable_node_histogram(ctx, 1) on every rank before tracing");
  if (capacity < s->n_node) return vsrt_fail(c, VSRT_E_CAPACITY, "the node histogram has %u entries", s->n_node);
  cudaSetDevice(c->device);
  CUDA_OK(c, cudaStreamSynchronize(s->rstream));
  CUDA_OK(c, cudaMemcpy(visits_of_slot, s->g_node.p, (size_t)s->n_node * 8, cudaMemcpyDeviceToHost));
  return VSRT_OK;
}

int vsrt_reduced_device(vsrt_context* c, void** counters_dev, void** treelet_hist_dev, uint64_t* n_treelets) {
  if (!c) return VSRT_E_INVALID;
  CommState* s = c->comm;
  if (!s) return vsrt_fail(c, VSRT_E_INVALID, "no communicator");
  if (counters_dev) *counters_dev = s->global;
  if (treelet_hist_dev) *treelet_hist_dev = s->g_hist.p;
  if (n_treelets) *n_treelets = s->n_hist == 0xFFFFFFFFu ? 0 : s->n_hist;
  return VSRT_OK;
}

}  // extern "C"
