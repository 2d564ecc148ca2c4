// vsrt_context.h -- the context behind the opaque vsrt_context of include/vsrt.h (the reference keeps this state in
// file-scope statics of vulkan_ray_tracing.cc: tlas_addr, blas_addr_map, treeletsFormed, rayCount, the treelet maps).
// Shared by vsrt_capi.cu (registration, formation, the K1 -> scan -> K3 pipeline) and reduce.cu (multi-GPU counter reduce).
#pragma once
#include "vsrt_internal.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct Reg { uint64_t host, size, dev; bool tlas; };

template <typename T> struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t n, bool keep = false, cudaStream_t st = nullptr) {
    if (n <= cap) return cudaSuccess;
    size_t nc = std::max(n, cap + cap / 2);
    T* q = nullptr; cudaError_t e = cudaMalloc(&q, nc * sizeof(T));
    if (e != cudaSuccess) { nc = n; e = cudaMalloc(&q, nc * sizeof(T)); if (e != cudaSuccess) return e; }
    if (keep && p && cap) cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (p) { if (keep) cudaDeviceSynchronize(); else cudaStreamSynchronize(st); cudaFree(p); }   // keep: copies on other streams may still read the old block
    p = q; cap = nc; return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Worker threads of the host-buffer calls (record expansion, treelet ids): started once per context, one job at a time; every
// worker runs job(t, n_threads) and the submitter carries on -- the frame pipeline launches the next window meanwhile.
struct HostPool {
  std::vector<std::thread> th; std::mutex m; std::condition_variable cv, cv_done;
  std::function<void(unsigned, unsigned)> job; uint64_t gen = 0; unsigned running = 0; bool stop = false;
  std::atomic<uint64_t> next{0};   // work cursor of the current job: the workers take pieces of it as they get to them (reset by submit)
  explicit HostPool(unsigned n) {
    for (unsigned t = 0; t < n; t++) th.emplace_back([this, t, n] {
      uint64_t seen = 0;
      for (;;) {
        std::function<void(unsigned, unsigned)> f;
        { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return stop || gen != seen; }); if (stop) return; seen = gen; f = job; }
        f(t, n);
        { std::lock_guard<std::mutex> l(m); if (--running == 0) cv_done.notify_all(); }
      }
    });
  }
  void wait() { std::unique_lock<std::mutex> l(m); cv_done.wait(l, [&] { return running == 0; }); }
  void submit(std::function<void(unsigned, unsigned)> f) { wait(); { std::lock_guard<std::mutex> l(m); job = std::move(f); running = (unsigned)th.size(); next.store(0, std::memory_order_relaxed); gen++; } cv.notify_all(); }
  ~HostPool() { wait(); { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); for (auto& x : th) x.join(); }
};
// Per-batch control words of a context, one device allocation: what a batch reads back (counters, record total, error flags) is
// its first DevCounters + 16 bytes -- one copy --, and one small kernel prepares all of it for a batch (counter backup for the
// rollback of a failed batch, flags, the ray counters of the queued traversal launches).
struct BatchCtl { DevCounters counters; unsigned long long total; uint32_t err; uint32_t sel; DevCounters bak; unsigned long long next_ray[4]; };
struct CommState;   // reduce.cu
struct HostStage { uint32_t* p = nullptr; uint64_t cap = 0; };   // pinned staging for one window's packed records (host-side expansion)

struct vsrt_context {
  vsrt_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // registration (allocTLAS / allocBLAS)
  std::vector<Reg> regs;
  bool committed = false;
  // arena
  uint8_t* d_arena = nullptr; uint64_t arena_bytes = 0;
  uint8_t* d_tarena = nullptr;   // K1's traversal copy of the arena, filled by treelet formation (same size, same slots)
  std::vector<Span> spans; Span* d_spans = nullptr;
  std::vector<BlasReg> blas; BlasReg* d_blas = nullptr;
  // treelets (treeletsFormed + the static maps)
  bool formed = false; uint64_t formed_tlas = 0; uint32_t formed_budget = 0;
  FormOutputs fo{}; FormResult fr{};
  std::vector<uint32_t> h_node_tid, h_tl_root; std::vector<uint64_t> h_tl_off, h_tl_node; bool mirrors = false;
  // per-batch buffers
  DevBuf<vsrt_ray> d_rays; DevBuf<vsrt_hit> d_hits; DevBuf<uint32_t> d_stage; DevBuf<uint32_t> d_counts;
  DevBuf<uint64_t> d_offsets; DevBuf<vsrt_txn> d_txns; DevBuf<uint32_t> d_tids; DevBuf<uint64_t> d_tid_addr; DevBuf<uint32_t> d_packed; DevBuf<uint8_t> d_scan_tmp;
  uint32_t stage_cap = 128;
  DevBuf<uint8_t> d_gstack;   // wavefront kernel: per-warp stack areas
  void* tb_tables = nullptr; DevBuf<uint8_t> d_tb; unsigned long long tb_stats[8] = { 0 };   // treelet-binned K1 (traverse_tb.cu): layout copy, scratch, statistics of the last batch
  DevBuf<uint32_t> d_nproc;   // procedural-leaf visits per ray
  BatchCtl* d_ctl = nullptr;   // the pointers below point into it
  DevCounters* d_counters = nullptr; DevCounters* d_counters_bak = nullptr; uint32_t* d_err = nullptr; uint32_t* d_sel = nullptr; unsigned long long* d_next_ray = nullptr;
  // pinned host memory: the per-batch read-backs (record total, error flags, counters) land here without a staging copy, and
  // small host-buffer calls (a warp's 32 rays) bounce their inputs and outputs through it so that a call is two queues of async
  // copies and two synchronisations instead of a blocking copy per array
  uint8_t* h_pin = nullptr;
  static constexpr size_t PIN_HEAD = 1024, PIN_BYTES = 4u << 20;
  DevCounters h_prev{};
  DevBuf<unsigned long long> d_hist; uint32_t hist_n = 0;
  bool node_hist_on = false; DevBuf<unsigned long long> d_node_hist; uint32_t node_hist_n = 0;   // optional per-slot visit counts
  // -remap_to_treelet_layout: where gpgpusim_malloc put treelet_layout_bvh (:1477), and the per-slot table
  uint64_t layout_base = 0; bool layout_base_set = false; DevBuf<uint64_t> d_remap; bool remap_valid = false;
  // replay helpers: sorted copy of the last trace, inverted treelet lists (slot -> (treelet, position))
  DevBuf<vsrt_txn> d_txns_sorted; DevBuf<uint32_t> d_tids_sorted; DevBuf<uint64_t> d_sort_keys;
  uint64_t* d_inv_off = nullptr; uint2* d_inv = nullptr;
  uint64_t order_min_rays = 65536;   // smaller batches keep the input order (the sort would cost more than it returns)
  DevBuf<uint8_t> d_order;    // rayorder.cu scratch: keys, sorted ray ids, decision word
  cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
  vsrt_device_results last{};
  uint64_t last_tlas = 0; int last_mode = 0; const vsrt_ray* last_rays = nullptr;
  bool last_packed_only = false; ArenaView last_av{};   // the last batch was delivered as packed records only (vsrt_trace_rays_packed)
  // window pipeline of the host-buffer calls: frame-level backup of counters / histograms, slot -> treelet root table, copy streams
  DevBuf<uint8_t> d_frame_bak; std::vector<uint64_t> h_root_of_slot; cudaStream_t copy_stream = nullptr, up_stream = nullptr; cudaEvent_t ev_copy[2] = { nullptr, nullptr }, ev_up[2] = { nullptr, nullptr }, ev_ready = nullptr;
  HostStage h_stage[2]; HostPool* pool = nullptr;
  CommState* comm = nullptr;   // multi-GPU reduce state (reduce.cu), NULL until vsrt_comm_init / vsrt_comm_attach
};

// shared helpers (vsrt_capi.cu)
int vsrt_fail(vsrt_context* c, int code, const char* fmt, ...);
void vsrt_comm_release(vsrt_context* c);          // reduce.cu: frees c->comm (called by vsrt_destroy)
int vsrt_comm_counters_reset(vsrt_context* c);     // reduce.cu: vsrt_reset_counters zeroes the reduce baselines and global totals too
void vsrt_comm_treelets_changed(vsrt_context* c); // reduce.cu: the treelet tables were re-formed, histograms restart
