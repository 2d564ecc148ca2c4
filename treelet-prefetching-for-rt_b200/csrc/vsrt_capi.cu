// vsrt_capi.cu -- host side of the C-ABI in include/vsrt.h: the context (the reference keeps this state in
// file-scope statics of vulkan_ray_tracing.cc: tlas_addr, blas_addr_map, treeletsFormed, rayCount, the treelet
// maps), arena upload, and the K1 -> scan -> K3 pipeline.  No CPU implementation of the path exists here: every
// entry point that computes launches the CUDA kernels of treelets.cu / traverse.cu / compact.cu.
#include "vsrt_context.h"
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <vector>

namespace {
thread_local char g_create_error[512] = "";
}  // namespace

int vsrt_fail(vsrt_context* c, int code, const char* fmt, ...) {
  char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf; else snprintf(g_create_error, sizeof(g_create_error), "%s", buf);
  return code;
}

namespace {

__global__ void k_add_base(unsigned long long* __restrict__ off, uint64_t n, unsigned long long base) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) off[i] += base;
}

// one launch instead of a device-to-device copy and three memsets: counter backup (restored if the batch fails), flags, record
// total, and the ray counters of the traversal launches queued for the batch (each launch pulls from its own)
__global__ void k_batch_prepare(BatchCtl* ctl) {
  const unsigned t = threadIdx.x;
  if (t < VSRT_COUNTERS_N_SUM + VSRT_COUNTERS_N_MAX) ctl->bak.v[t] = ctl->counters.v[t];
  if (t < 4) ctl->next_ray[t] = 0ull;
  if (t == 4) { ctl->err = 0u; ctl->total = 0ull; }
}

template <typename... A> int fail(vsrt_context* c, int code, const char* fmt, A... a) { return vsrt_fail(c, code, fmt, a...); }
#define CUDA_OK(c, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(c, VSRT_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); } while (0)

void free_treelets(vsrt_context* c) {
  cudaFree(c->fo.node_tid); cudaFree(c->fo.root_bits); cudaFree(c->fo.root_prefix); cudaFree(c->fo.tl_root); cudaFree(c->fo.tl_off); cudaFree(c->fo.tl_node); cudaFree(c->fo.hot_keys);
  c->fo = FormOutputs{}; c->formed = false; c->mirrors = false; c->hist_n = 0; c->remap_valid = false;
  cudaFree(c->d_inv_off); cudaFree(c->d_inv); c->d_inv_off = nullptr; c->d_inv = nullptr;
  c->h_node_tid.clear(); c->h_tl_root.clear(); c->h_tl_off.clear(); c->h_tl_node.clear(); c->h_root_of_slot.clear();
  vsrt_comm_treelets_changed(c);
  vsrt_tb_free_layout(c->tb_tables); c->tb_tables = nullptr;
}

const Reg* find_tlas(const vsrt_context* c, uint64_t host) {
  const Reg* r = nullptr;
  for (const Reg& x : c->regs) if (x.tlas && x.host == host) r = &x;   // the last registration wins, like tlas_addr (:4896)
  return r;
}
bool host_to_slot_h(const vsrt_context* c, uint64_t host, uint32_t* slot) {
  for (const Span& s : c->spans) if (host >= s.host && host - s.host < s.size) { if ((host - s.host) & 63) return false; *slot = s.slot0 + (uint32_t)((host - s.host) >> 6); return true; }
  return false;
}
uint64_t slot_to_host_h(const vsrt_context* c, uint32_t slot) {
  for (size_t i = c->spans.size(); i-- > 0;) if (c->spans[i].slot0 <= slot) return c->spans[i].host + (uint64_t)(slot - c->spans[i].slot0) * 64;
  return 0;
}

int make_view(vsrt_context* c, uint64_t tlas_host, ArenaView* av) {
  const Reg* t = find_tlas(c, tlas_host);
  if (!t) return fail(c, VSRT_E_UNKNOWN_AS, "TLAS %p was never registered with vsrt_alloc_tlas (reference: abort(), vulkan_ray_tracing.cc:1568)", (void*)tlas_host);
  uint32_t slot = 0;
  if (!host_to_slot_h(c, tlas_host, &slot)) return fail(c, VSRT_E_UNKNOWN_AS, "TLAS address not inside the committed arena");
  av->base = c->d_arena; av->n_slots = (uint32_t)(c->arena_bytes / 64); av->n_spans = (uint32_t)c->spans.size(); av->n_blas = (uint32_t)c->blas.size();
  av->tlas_slot = slot; av->spans = c->d_spans; av->blas = c->d_blas; av->tlas_delta = (int64_t)(t->dev - t->host);
  av->uniform_delta = 1; av->force_exact = (c->formed && c->fr.nonfinite) ? 1u : 0u;
  av->inst_base = c->formed ? c->fr.inst_base : 0u; av->pad2 = 0;
  for (const BlasReg& b : c->blas) if (b.delta != av->tlas_delta) av->uniform_delta = 0;
  return VSRT_OK;
}
TreeletView treelet_view(const vsrt_context* c) {
  TreeletView tv; tv.node_tid = c->fo.node_tid; tv.root_bits = c->fo.root_bits; tv.root_prefix = c->fo.root_prefix; tv.tl_root = c->fo.tl_root;
  tv.n_treelets = c->fr.n_treelets; tv.pad = 0; tv.tnodes = c->fo.tnodes; tv.hot_keys = c->fo.hot_keys; return tv;
}

int ensure_mirrors(vsrt_context* c) {
  if (c->mirrors) return VSRT_OK;
  const size_t ns = c->arena_bytes / 64, nt = c->fr.n_treelets, ne = c->fr.n_entries;
  c->h_node_tid.resize(ns); c->h_tl_root.resize(nt); c->h_tl_off.resize(nt + 1); c->h_tl_node.resize(ne);
  CUDA_OK(c, cudaMemcpy(c->h_node_tid.data(), c->fo.node_tid, ns * 4, cudaMemcpyDeviceToHost));
  for (uint32_t& t : c->h_node_tid) if (t != VSRT_NO_TID) t &= VSRT_TID_MASK;
  CUDA_OK(c, cudaMemcpy(c->h_tl_root.data(), c->fo.tl_root, nt * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(c, cudaMemcpy(c->h_tl_off.data(), c->fo.tl_off, (nt + 1) * 8, cudaMemcpyDeviceToHost));
  if (ne) CUDA_OK(c, cudaMemcpy(c->h_tl_node.data(), c->fo.tl_node, ne * 8, cudaMemcpyDeviceToHost));
  c->mirrors = true; return VSRT_OK;
}
// device address of a list entry: BLAS headers carry their own allocBLAS offset (:1149-1153), everything else the TLAS offset
uint64_t entry_dev_addr(const vsrt_context* c, uint64_t entry, int64_t tlas_delta) {
  const uint32_t slot = (uint32_t)entry, kind = (uint32_t)(entry >> 32);
  const uint64_t host = slot_to_host_h(c, slot);
  if (kind == K_BLAS_HEADER) for (const BlasReg& b : c->blas) if (b.hdr_slot == slot) return host + (uint64_t)b.delta;
  return host + (uint64_t)tlas_delta;
}
int64_t formed_delta(const vsrt_context* c) { const Reg* t = find_tlas(c, c->formed_tlas); return t ? (int64_t)(t->dev - t->host) : 0; }

// simulated-device address -> slot, honouring that BLAS headers are keyed by their allocBLAS address
bool dev_addr_to_slot(const vsrt_context* c, uint64_t addr, uint32_t* slot) {
  for (const BlasReg& b : c->blas) if (slot_to_host_h(c, b.hdr_slot) + (uint64_t)b.delta == addr) { *slot = b.hdr_slot; return true; }
  return host_to_slot_h(c, addr - (uint64_t)formed_delta(c), slot);
}

int do_form(vsrt_context* c, uint64_t tlas, uint32_t budget) {
  int rc = vsrt_commit(c); if (rc) return rc;
  if (c->formed && c->formed_tlas == tlas && c->formed_budget == budget) return VSRT_OK;
  free_treelets(c);
  ArenaView av; rc = make_view(c, tlas, &av); if (rc) return rc;
  char eb[400] = "";
  rc = vsrt_launch_form_treelets(av, budget, c->stream, &c->fo, &c->fr, c->d_err, eb, sizeof(eb), c->d_tarena);
  if (rc) return fail(c, rc, "%s", eb);
  c->formed = true; c->formed_tlas = tlas; c->formed_budget = budget;
  CUDA_OK(c, c->d_hist.ensure(std::max<size_t>(c->fr.n_treelets, 1)));
  CUDA_OK(c, cudaMemsetAsync(c->d_hist.p, 0, (size_t)c->fr.n_treelets * 8, c->stream));
  c->hist_n = c->fr.n_treelets;
  // a trace call may run on a caller stream: everything queued on the context's stream above must have landed first
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VSRT_OK;
}

// packed_out: K3 writes the 4-byte packed records into d_packed instead of the 16-byte records + treelet indices (the host
// form of vsrt_trace_rays_packed); ensure_full_records() expands them later if a caller asks for the full form after all
// A frame may be traced as several WINDOWS (chunks of consecutive rays, for the pipelined host calls): d_rays / n are the window,
// r0 its first ray in the frame, frame_n the frame's ray count (0 = the window is the frame) and rec_base the records of the
// windows before it.  Every per-ray and per-record buffer is frame-sized and a window writes its own slice, so after the last
// window the context holds the whole frame exactly as a single batch would have left it.
int run_batch(vsrt_context* c, uint64_t tlas, int mode, const vsrt_ray* d_rays, uint64_t n, cudaStream_t st, bool packed_out = false,
              uint64_t r0 = 0, uint64_t frame_n = 0, uint64_t rec_base = 0) {
  if (mode != VSRT_MODE_DFS && mode != VSRT_MODE_TREELET) return fail(c, VSRT_E_INVALID, "mode must be VSRT_MODE_DFS or VSRT_MODE_TREELET");
  if (n >= (1ull << 32) - 1) return fail(c, VSRT_E_INVALID, "a batch holds at most 2^32-2 rays; split the frame");
  int rc = do_form(c, tlas, c->cfg.max_treelet_size); if (rc) return rc;   // lazily, like :1593 / :2364
  const uint64_t remap_pitch = c->cfg.remap_to_treelet_layout ? (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride : 0;
  if (c->cfg.remap_to_treelet_layout) {
    if (!c->layout_base_set) return fail(c, VSRT_E_INVALID, "remap_to_treelet_layout needs vsrt_set_treelet_layout_base (the address gpgpusim_malloc gave treelet_layout_bvh, vulkan_ray_tracing.cc:1477)");
    if (!c->remap_valid) {
      const uint32_t ns = (uint32_t)(c->arena_bytes / 64);
      CUDA_OK(c, c->d_remap.ensure(std::max<uint32_t>(ns, 1)));
      rc = vsrt_launch_remap(c->fo, c->fr.n_treelets, ns, c->layout_base, remap_pitch, c->d_remap.p, c->stream);
      if (rc) return fail(c, rc, "treelet-layout remap kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
      c->remap_valid = true;
    }
  }
  ArenaView av; rc = make_view(c, tlas, &av); if (rc) return rc;
  const TreeletView tv = treelet_view(c);
  const uint64_t fn = frame_n ? frame_n : n;
  const vsrt_device_results before = c->last;
  if (r0 == 0) c->last = vsrt_device_results{};
  c->last_tlas = tlas; c->last_mode = mode; c->last_rays = d_rays - r0;
  CUDA_OK(c, c->d_hits.ensure(std::max<uint64_t>(fn, 1))); CUDA_OK(c, c->d_counts.ensure(std::max<uint64_t>(fn, 1))); CUDA_OK(c, c->d_offsets.ensure(fn + 1));
  CUDA_OK(c, c->d_scan_tmp.ensure(vsrt_scan_tmp_bytes(n))); CUDA_OK(c, c->d_nproc.ensure(std::max<uint64_t>(fn, 1)));
  uint32_t launches = 0; uint64_t total = 0;
  for (int attempt = 0;; attempt++) {
    CUDA_OK(c, c->d_stage.ensure(std::max<uint64_t>(fn, 1) * c->stage_cap));
    // this window's slices of the frame-sized buffers
    vsrt_hit* const w_hits = c->d_hits.p + r0; uint32_t* const w_counts = c->d_counts.p + r0; uint64_t* const w_offsets = c->d_offsets.p + r0;
    uint32_t* const w_nproc = c->d_nproc.p + r0; uint32_t* const w_stage = c->d_stage.p + r0 * c->stage_cap;
    k_batch_prepare<<<1, 32, 0, st>>>(c->d_ctl);
    TraverseParams tp; tp.av = av; tp.tv = tv; tp.rays = d_rays; tp.n_rays = n; tp.hits = w_hits; tp.stage = w_stage; tp.counts = w_counts; tp.nproc = w_nproc;
    tp.cap = c->stage_cap; tp.mode = (uint32_t)mode; tp.counters = c->d_counters; tp.err_flags = c->d_err; tp.next_ray = c->d_next_ray;
    { const char* a = getenv("VSRT_REFILL_T"); const char* b = getenv("VSRT_LEAF_T"); tp.refill_t = a ? (uint32_t)atoi(a) : 8u; tp.leaf_t = b ? (uint32_t)atoi(b) : 1u; }   // round-2 sweep: leaf threshold 1 is best on the headline (1 %) and on the incoherent configs (6 %)
    tp.magic16 = 0x64646464u; tp.only_deferred = 0; tp.gate = 0; tp.perm = nullptr; tp.perm_on = nullptr; tp.sel = nullptr; tp.sel_want = 0;
    // ray order (rayorder.cu): which rays share a warp; batches too small to fill the GPU twice are left alone
    uint32_t ray_order = c->cfg.ray_order;
    if (const char* ro = getenv("VSRT_RAY_ORDER")) ray_order = (uint32_t)atoi(ro);
    // AUTO is the input order: on every workload measured (camera rays, pixel-ordered bounce rays, a triangle soup's bounce rays,
    // uniformly random rays) the sorted order raised K1's L1 / L2 hit rates and cut its DRAM reads by up to 30 %, and still made it
    // 15-30 % SLOWER (profiles/README.md, round 2) -- so sorting is something a caller asks for, not something the library guesses
    const bool want_order = ray_order == VSRT_RAY_ORDER_SORTED && n >= c->order_min_rays;
    CUDA_OK(c, cudaEventRecord(c->ev[4], st));
    if (want_order) {
      CUDA_OK(c, c->d_order.ensure(vsrt_rayorder_tmp_bytes(n)));
      rc = vsrt_launch_rayorder(d_rays, n, true, c->d_order.p, &tp.perm, &tp.perm_on, st);
      if (rc) return fail(c, rc, "ray-order kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
      launches += 8;   // bounds, keys, 3 x (count, scan, scatter) -- approximate
    }
    const uint32_t stack_entries = c->cfg.stack_entries ? c->cfg.stack_entries : 96;
    // K1 variant: the lane-owned kernel (traverse.cu) is the default; VSRT_K1_WF=1 selects the warp-wavefront kernel
    // (traverse_wf.cu), bit-identical results, measured 10 % slower on the bench workload (profiles/README.md)
    const bool wavefront = getenv("VSRT_K1_WF") && atoi(getenv("VSRT_K1_WF")) != 0;
    // VSRT_K1_TB=1: the treelet-binned wavefront kernel (traverse_tb.cu; traceRayWithTreelets only): rounds of "bin the rays by next
    // treelet, stage shared treelets with TMA, drain"; bit-identical results, measured slower (profiles/README.md)
    const bool binned = !wavefront && mode == VSRT_MODE_TREELET && !av.force_exact && n && getenv("VSRT_K1_TB") && atoi(getenv("VSRT_K1_TB")) != 0;
    if (binned) {
      if (!c->tb_tables) { rc = vsrt_tb_build_layout(av, c->fo, c->fr.n_treelets, &c->tb_tables, st); if (rc) return fail(c, rc, "building the treelet-layout copy failed: %s", cudaGetErrorString(cudaGetLastError())); }
      CUDA_OK(c, c->d_tb.ensure(vsrt_tb_scratch_bytes(n, stack_entries)));
    }
    unsigned wf_grid = 0;
    if (wavefront && n) {
      wf_grid = vsrt_wf_grid(n);
      CUDA_OK(c, c->d_gstack.ensure(vsrt_wf_stack_bytes(wf_grid, stack_entries)));
      tp.gstack = (uint2*)c->d_gstack.p; tp.stack_n = stack_entries;
    } else if (const size_t gb = vsrt_traverse_gstack_bytes(stack_entries)) {   // VSRT_K1_GSTACK build: the hot kernel's stack lives in global memory
      CUDA_OK(c, c->d_gstack.ensure(gb));
      tp.gstack = (uint2*)c->d_gstack.p; tp.stack_n = stack_entries;
    } else { tp.gstack = nullptr; tp.stack_n = stack_entries; }
    CUDA_OK(c, cudaEventRecord(c->ev[0], st));
    if (binned) { tp.perm = nullptr; tp.perm_on = nullptr; rc = vsrt_launch_traverse_tb(tp, c->tb_tables, stack_entries, c->d_tb.p, c->tb_stats, st); }
    else if (wavefront) rc = vsrt_launch_traverse_wf(tp, wf_grid, av.force_exact != 0, st);
    else if (av.force_exact) rc = vsrt_launch_traverse(tp, stack_entries, true, false, st);
    else {
      // Which node layout the hot kernel reads: the traversal copy saves the byte shuffles of every node visit and
      // wins where K1 is issue-bound (camera rays: 1.9 %); where it waits for load latency (bounce rays) the Mesa-layout kernel is
      // 3-8 % faster.  A frame-sized batch is therefore sampled on the device (256 pairs of consecutive rays: do they point the same
      // way?) and BOTH instantiations are queued, each gated on the sample's verdict -- no read-back; the one that is not picked
      // returns at once.  VSRT_K1_LAYOUT = 0 / 1 forces the Mesa layout / the traversal copy; small batches take the traversal copy.
      int layout = -1;
      if (const char* e = getenv("VSRT_K1_LAYOUT")) layout = atoi(e);
      if (layout < 0 && n >= 65536) {
        rc = vsrt_launch_ray_coherence(d_rays, n, c->d_sel, st);
        tp.sel = c->d_sel; tp.sel_want = 1;
        if (!rc) rc = vsrt_launch_traverse(tp, stack_entries, false, true, st);
        tp.sel_want = 0; tp.next_ray = c->d_next_ray + 1;
        if (!rc) rc = vsrt_launch_traverse(tp, stack_entries, false, false, st);
        tp.sel = nullptr; tp.next_ray = c->d_next_ray;
        launches += 2;
      } else rc = vsrt_launch_traverse(tp, stack_entries, false, layout != 0, st);
    }
    if (rc) return fail(c, rc, "traversal kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    launches += n ? 1 : 0;
    if (!av.force_exact && n) {
      // rays (or instances) with non-finite coordinates are deferred by the fast kernel (EF_NEED_EXACT): the EXACT kernel is
      // queued behind it and returns at once when nothing was deferred -- no flag read-back between the two
      tp.only_deferred = 1; tp.gate = EF_NEED_EXACT; tp.next_ray = c->d_next_ray + 2;
      rc = wavefront ? vsrt_launch_traverse_wf(tp, wf_grid, true, st) : vsrt_launch_traverse(tp, stack_entries, true, false, st);
      if (rc) return fail(c, rc, "exact traversal kernel launch failed");
      launches++;
    }
    CUDA_OK(c, cudaEventRecord(c->ev[1], st));
    rc = vsrt_launch_scan(w_counts, n, w_offsets, c->d_scan_tmp.p, st, &c->d_ctl->total); if (rc) return fail(c, rc, "scan launch failed");
    CUDA_OK(c, cudaEventRecord(c->ev[2], st));
    launches += n ? 1 : 0;   // the one-pass scan
    bool node_hist_queued = false;
    auto queue_node_hist = [&]() -> int {
      if (!c->node_hist_on || !n || node_hist_queued) return VSRT_OK;
      const uint32_t ns = (uint32_t)(c->arena_bytes / 64);
      if (c->node_hist_n != ns) { if (c->d_node_hist.ensure(std::max<uint32_t>(ns, 1)) != cudaSuccess || cudaMemsetAsync(c->d_node_hist.p, 0, (size_t)ns * 8, st) != cudaSuccess) return VSRT_E_CUDA; c->node_hist_n = ns; }
      node_hist_queued = true; launches++;
      return vsrt_launch_node_hist(w_stage, c->stage_cap, w_offsets, n, c->d_node_hist.p, c->d_err, EF_BAD_BVH | EF_STACK | EF_TRACE_CAP | EF_UNSUPPORTED, st);
    };
    rc = queue_node_hist(); if (rc) return fail(c, rc, "node histogram launch failed");
    // K3 is queued right away into the buffers of the previous batch; it checks the error flags and the record count on the
    // device and does nothing if either says no.  One host synchronisation per batch in the steady state.
    CompactParams cp; cp.av = av; cp.tv = tv; cp.stage = w_stage; cp.cap = c->stage_cap; cp.mode = (uint32_t)mode; cp.offsets = w_offsets; cp.n_rays = n;
    cp.counters = c->d_counters; cp.treelet_hist = getenv("VSRT_NO_HIST") ? nullptr : c->d_hist.p;
    cp.remap = c->cfg.remap_to_treelet_layout ? c->d_remap.p : nullptr;
    cp.err_flags = c->d_err; cp.fatal_mask = EF_BAD_BVH | EF_STACK | EF_TRACE_CAP | EF_UNSUPPORTED;
    cp.packed = nullptr; cp.count = 1; cp.pad3 = 0;
    const uint64_t have_cap = packed_out ? c->d_packed.cap : std::min(c->d_txns.cap, c->d_tids.cap);
    uint64_t queued_cap = have_cap > rec_base ? have_cap - rec_base : 0;
    if (queued_cap && n) {
      cp.txns = c->d_txns.p + rec_base; cp.tids = c->d_tids.p + rec_base; cp.packed = packed_out ? c->d_packed.p + rec_base : nullptr; cp.out_capacity = queued_cap;
      rc = vsrt_launch_compact(cp, st); if (rc) return fail(c, rc, "compaction kernel launch failed");
      launches++;
    }
    CUDA_OK(c, cudaEventRecord(c->ev[3], st));
    uint32_t h_err = 0; DevCounters now;
    static_assert(16 + sizeof(DevCounters) <= vsrt_context::PIN_HEAD, "read-back block must fit the head of the pinned buffer");
    static_assert(offsetof(BatchCtl, total) == sizeof(DevCounters) && offsetof(BatchCtl, err) == sizeof(DevCounters) + 8, "read-back block layout");
    CUDA_OK(c, cudaMemcpyAsync(c->h_pin, c->d_ctl, sizeof(DevCounters) + 16, cudaMemcpyDeviceToHost, st));     // counters | record total (the scan left it there) | error flags
    CUDA_OK(c, cudaStreamSynchronize(st));
    memcpy(&now, c->h_pin, sizeof(now)); memcpy(&total, c->h_pin + sizeof(DevCounters), 8); memcpy(&h_err, c->h_pin + sizeof(DevCounters) + 8, 4);
    if (h_err & (EF_BAD_BVH | EF_STACK | EF_UNSUPPORTED)) {
      // a failed batch leaves no trace in the counters (rayCount, g_rt_* and accessedDataSize are restored); K3 did not run
      cudaMemcpyAsync(c->d_counters, c->d_counters_bak, sizeof(DevCounters), cudaMemcpyDeviceToDevice, st);
      cudaStreamSynchronize(st);
      c->last = vsrt_device_results{};
      (void)before;
      if (h_err & EF_BAD_BVH) return fail(c, VSRT_E_BAD_BVH, "traversal met a malformed node");
      if (h_err & EF_STACK) return fail(c, VSRT_E_STACK_OVERFLOW, "a ray needed more than %u traversal-stack entries; raise vsrt_config.stack_entries (at most 384)", c->cfg.stack_entries ? c->cfg.stack_entries : 96);
      return fail(c, VSRT_E_UNSUPPORTED, "a ray visited more than 4095 procedural leaves (or 2^20 - 1 nodes): beyond what the per-ray staging segment records");
    }
    if (h_err & EF_TRACE_CAP) {   // a ray produced more records than its staging segment holds: grow and redo the batch
      if (attempt >= 8) return fail(c, VSRT_E_CAPACITY, "per-ray trace staging overflow");
      CUDA_OK(c, cudaMemcpyAsync(c->d_counters, c->d_counters_bak, sizeof(DevCounters), cudaMemcpyDeviceToDevice, st));
      c->stage_cap *= 2;
      continue;
    }
    if (!queued_cap || total > queued_cap) {
      // first batch, or more records than the buffers held: the queued K3 declined; grow and run it now
      // (the first window of a chunked frame sizes the buffers for the whole frame from its own records per ray)
      uint64_t need = rec_base + total;
      if (r0 == 0 && fn > n && n) need = (uint64_t)((double)total * ((double)fn / (double)n) * 1.15) + 4096;
      need = std::max<uint64_t>(need, 1);
      if (packed_out) { CUDA_OK(c, c->d_packed.ensure(need, rec_base > 0, st)); cp.packed = c->d_packed.p + rec_base; cp.out_capacity = c->d_packed.cap - rec_base; }
      else {
        CUDA_OK(c, c->d_txns.ensure(need, rec_base > 0, st)); CUDA_OK(c, c->d_tids.ensure(need, rec_base > 0, st));
        cp.txns = c->d_txns.p + rec_base; cp.tids = c->d_tids.p + rec_base; cp.out_capacity = std::min(c->d_txns.cap, c->d_tids.cap) - rec_base;
      }
      if (n) {
        rc = vsrt_launch_compact(cp, st); if (rc) return fail(c, rc, "compaction kernel launch failed");
        launches++;
      }
      CUDA_OK(c, cudaEventRecord(c->ev[3], st));
      CUDA_OK(c, cudaMemcpyAsync(c->h_pin, c->d_counters, sizeof(now), cudaMemcpyDeviceToHost, st));
      CUDA_OK(c, cudaStreamSynchronize(st));
      memcpy(&now, c->h_pin, sizeof(now));
    }
    // rayCount (:1665) advanced by the traversal kernel; accessedDataSize delta of this batch = its algorithmic bytes
    c->last.algorithmic_bytes = (r0 ? before.algorithmic_bytes : 0) + now.v[CI_ACCESSED] - c->h_prev.v[CI_ACCESSED];
    c->h_prev = now;
    // a later window's offsets continue where the earlier ones ended (K3 and the node histogram have read them window-relative)
    if (rec_base && n) k_add_base<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>((unsigned long long*)(c->d_offsets.p + r0), n + 1, rec_base);
    break;
  }
  c->last.hits = c->d_hits.p; c->last.trace_offsets = c->d_offsets.p;
  c->last.txns = packed_out ? nullptr : c->d_txns.p; c->last.treelet_ids = packed_out ? nullptr : c->d_tids.p;
  c->last_packed_only = packed_out; c->last_av = av;
  c->last.n_rays = r0 + n; c->last.n_txn = rec_base + total; c->last.kernel_launches = (r0 ? before.kernel_launches : 0) + launches;
  float t_order = 0, t_trav = 0, t_scan = 0, t_comp = 0;
  cudaEventElapsedTime(&t_order, c->ev[4], c->ev[0]); cudaEventElapsedTime(&t_trav, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&t_scan, c->ev[1], c->ev[2]); cudaEventElapsedTime(&t_comp, c->ev[2], c->ev[3]);
  c->last.order_ms = (r0 ? before.order_ms : 0.0f) + t_order; c->last.traverse_ms = (r0 ? before.traverse_ms : 0.0f) + t_trav;
  c->last.scan_ms = (r0 ? before.scan_ms : 0.0f) + t_scan; c->last.compact_ms = (r0 ? before.compact_ms : 0.0f) + t_comp;
  return VSRT_OK;
}

int trace_frame_pipelined(vsrt_context* c, uint64_t tlas, int mode, uint64_t n, const vsrt_ray* rays, vsrt_hit* hits, uint64_t* trace_offsets,
                          bool packed, void* records_out, uint64_t capacity, uint64_t* treelet_ids, uint64_t* n_txn, uint64_t chunk);

// The last batch was delivered in packed form only: expand its staged records into the 16-byte records + treelet indices now
// (same K3, counters and histogram untouched -- they were accumulated when the batch ran).
int ensure_full_records(vsrt_context* c) {
  if (!c->last_packed_only) return VSRT_OK;
  const uint64_t total = c->last.n_txn, n = c->last.n_rays;
  CUDA_OK(c, c->d_txns.ensure(std::max<uint64_t>(total, 1))); CUDA_OK(c, c->d_tids.ensure(std::max<uint64_t>(total, 1)));
  if (n) {
    CompactParams cp; cp.av = c->last_av; cp.tv = treelet_view(c); cp.stage = c->d_stage.p; cp.cap = c->stage_cap; cp.mode = (uint32_t)c->last_mode;
    cp.offsets = c->d_offsets.p; cp.n_rays = n; cp.counters = c->d_counters; cp.treelet_hist = nullptr;
    cp.remap = c->cfg.remap_to_treelet_layout ? c->d_remap.p : nullptr; cp.err_flags = nullptr; cp.fatal_mask = 0;
    cp.txns = c->d_txns.p; cp.tids = c->d_tids.p; cp.out_capacity = std::min(c->d_txns.cap, c->d_tids.cap); cp.packed = nullptr; cp.count = 0; cp.pad3 = 0;
    int rc = vsrt_launch_compact(cp, c->stream); if (rc) return fail(c, rc, "compaction kernel launch failed");
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  c->last.txns = c->d_txns.p; c->last.treelet_ids = c->d_tids.p; c->last_packed_only = false;
  return VSRT_OK;
}

}  // namespace

// ================================================================= C-ABI
extern "C" {

void vsrt_default_config(vsrt_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->device = -1; cfg->max_treelet_size = 49152;   // gpu-sim.cc:910-912
  cfg->stack_entries = 96;
}

int vsrt_config_parse(vsrt_config* cfg, const char* text) {
  if (!cfg || !text) return VSRT_E_INVALID;
  const char* p = text;
  while (*p) {
    const char* eol = strchr(p, '\n'); size_t len = eol ? (size_t)(eol - p) : strlen(p);
    std::string line(p, len); p += len + (eol ? 1 : 0);
    size_t h = line.find('#'); if (h != std::string::npos) line.resize(h);
    char name[128]; long long val;
    if (sscanf(line.c_str(), " -%127s %lld", name, &val) == 2) {
      if (!strcmp(name, "max_treelet_size")) cfg->max_treelet_size = (uint32_t)val;
      else if (!strcmp(name, "treelet_based_traversal")) cfg->treelet_based_traversal = (uint32_t)val;
      else if (!strcmp(name, "remap_to_treelet_layout")) cfg->remap_to_treelet_layout = (uint32_t)val;
      else if (!strcmp(name, "treelet_remap_stride")) cfg->treelet_remap_stride = (uint32_t)val;
      else if (!strcmp(name, "load_treelet_metadata")) cfg->load_treelet_metadata = (uint32_t)val;
    }
  }
  return VSRT_OK;
}

int vsrt_create(const vsrt_config* cfg, vsrt_context** out) {
  if (!out) return VSRT_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(nullptr, VSRT_E_NO_DEVICE, "no CUDA device (%s); libvsrt has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  vsrt_context* c = new vsrt_context();
  if (cfg) c->cfg = *cfg; else vsrt_default_config(&c->cfg);
  if (c->cfg.device >= 0) { if (c->cfg.device >= ndev || cudaSetDevice(c->cfg.device) != cudaSuccess) { delete c; return fail(nullptr, VSRT_E_NO_DEVICE, "cannot select CUDA device %d", cfg->device); } }
  cudaGetDevice(&c->device);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, c->device);
  if (prop.major < 10) { delete c; return fail(nullptr, VSRT_E_NO_DEVICE, "device %d is sm_%d%d; libvsrt is built for sm_100a only", c->device, prop.major, prop.minor); }
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_ctl, sizeof(BatchCtl)) == cudaSuccess && cudaMemset(c->d_ctl, 0, sizeof(BatchCtl)) == cudaSuccess;
  if (ok) { c->d_counters = &c->d_ctl->counters; c->d_counters_bak = &c->d_ctl->bak; c->d_err = &c->d_ctl->err; c->d_sel = &c->d_ctl->sel; c->d_next_ray = c->d_ctl->next_ray; }
  ok = ok && cudaMallocHost(&c->h_pin, vsrt_context::PIN_BYTES) == cudaSuccess;
  for (int i = 0; i < 5 && ok; i++) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  if (!ok) { const char* m = cudaGetErrorString(cudaGetLastError()); vsrt_destroy(c); return fail(nullptr, VSRT_E_NO_DEVICE, "CUDA initialisation failed: %s", m); }
  if (c->cfg.max_treelet_size == 0) c->cfg.max_treelet_size = 49152;
  if (c->cfg.stack_entries > 384) { vsrt_destroy(c); return fail(nullptr, VSRT_E_INVALID, "vsrt_config.stack_entries = %u: the traversal kernel is built for at most 384 entries per ray", cfg->stack_entries); }
  // initial staging records per ray (doubles, with the batch redone, whenever a ray outgrows it); the knob exists for the tests
  if (const char* om = getenv("VSRT_RAY_ORDER_MIN")) { const long long v = atoll(om); if (v >= 1) c->order_min_rays = (uint64_t)v; }
  if (const char* sc = getenv("VSRT_STAGE_CAP")) { const int v = atoi(sc); if (v >= 4 && v <= (1 << 20)) c->stage_cap = (uint32_t)v; }
  *out = c;
  return VSRT_OK;
}

void vsrt_destroy(vsrt_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  vsrt_comm_release(c);
  free_treelets(c);
  cudaFree(c->d_arena); cudaFree(c->d_tarena); cudaFree(c->d_spans); cudaFree(c->d_blas); cudaFree(c->d_ctl);
  c->d_rays.release(); c->d_hits.release(); c->d_gstack.release(); c->d_nproc.release(); c->d_stage.release(); c->d_counts.release(); c->d_offsets.release(); c->d_txns.release();
  c->d_tids.release(); c->d_tid_addr.release(); c->d_packed.release(); c->d_scan_tmp.release(); c->d_hist.release(); c->d_remap.release();
  c->d_txns_sorted.release(); c->d_tids_sorted.release(); c->d_sort_keys.release(); c->d_order.release(); c->d_tb.release(); c->d_node_hist.release(); c->d_frame_bak.release();
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream); if (c->up_stream) cudaStreamDestroy(c->up_stream);
  for (int i = 0; i < 2; i++) { if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]); if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]); }
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  delete c->pool; c->pool = nullptr;
  for (HostStage& hs : c->h_stage) if (hs.p) cudaFreeHost(hs.p);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* vsrt_last_error(const vsrt_context* c) { return c ? c->err.c_str() : g_create_error; }

static int add_reg(vsrt_context* c, const void* root, uint64_t size, uint64_t dev, bool tlas) {
  if (!c || !root || size < 64) return c ? fail(c, VSRT_E_INVALID, "registration needs a non-null address and at least one 64-byte record") : VSRT_E_INVALID;
  if (((uintptr_t)root & 63) != 0) return fail(c, VSRT_E_INVALID, "AS buffer %p is not 64-byte aligned", root);
  for (Reg& r : c->regs) if (r.host == (uint64_t)(uintptr_t)root && r.tlas == tlas) { r.size = size; r.dev = dev; c->committed = false; return VSRT_OK; }   // std::map overwrite
  c->regs.push_back(Reg{ (uint64_t)(uintptr_t)root, size, dev, tlas });
  c->committed = false;
  return VSRT_OK;
}
int vsrt_alloc_tlas(vsrt_context* c, const void* root, uint64_t size, uint64_t dev) { return add_reg(c, root, size, dev, true); }
int vsrt_alloc_blas(vsrt_context* c, const void* root, uint64_t size, uint64_t dev) { return add_reg(c, root, size, dev, false); }

int vsrt_commit(vsrt_context* c) {
  if (!c) return VSRT_E_INVALID;
  if (c->committed) return VSRT_OK;
  if (c->regs.empty()) return fail(c, VSRT_E_UNKNOWN_AS, "no acceleration structure registered");
  cudaSetDevice(c->device);
  free_treelets(c);
  // merge the registered ranges into disjoint spans, ascending by host address
  std::vector<std::pair<uint64_t, uint64_t>> iv;
  for (const Reg& r : c->regs) iv.push_back({ r.host, r.host + ((r.size + 63) & ~63ull) });
  std::sort(iv.begin(), iv.end());
  std::vector<std::pair<uint64_t, uint64_t>> mg;
  for (auto& x : iv) { if (!mg.empty() && x.first <= mg.back().second) mg.back().second = std::max(mg.back().second, x.second); else mg.push_back(x); }
  c->spans.clear(); uint64_t slots = 0;
  for (auto& x : mg) { Span s; s.host = x.first; s.size = x.second - x.first; s.slot0 = (uint32_t)slots; s.n_slots = (uint32_t)(s.size / 64); c->spans.push_back(s); slots += s.size / 64; }
  if (slots >= (1ull << 29)) return fail(c, VSRT_E_UNSUPPORTED, "arena of %llu bytes exceeds the 32 GiB the 29-bit trace record addresses", (unsigned long long)(slots * 64));
  cudaFree(c->d_arena); cudaFree(c->d_tarena); cudaFree(c->d_spans); cudaFree(c->d_blas); c->d_arena = nullptr; c->d_tarena = nullptr; c->d_spans = nullptr; c->d_blas = nullptr;
  c->arena_bytes = slots * 64;
  CUDA_OK(c, cudaMalloc(&c->d_arena, c->arena_bytes));
  CUDA_OK(c, cudaMalloc(&c->d_tarena, c->arena_bytes));   // K1's traversal copy (filled by formation)
  for (const Span& s : c->spans) CUDA_OK(c, cudaMemcpyAsync(c->d_arena + (uint64_t)s.slot0 * 64, (const void*)(uintptr_t)s.host, s.size, cudaMemcpyHostToDevice, c->stream));
  c->blas.clear();
  for (const Reg& r : c->regs) if (!r.tlas) { BlasReg b; uint32_t slot = 0; host_to_slot_h(c, r.host, &slot); b.hdr_slot = slot; b.pad = 0; b.delta = (int64_t)(r.dev - r.host); c->blas.push_back(b); }
  std::sort(c->blas.begin(), c->blas.end(), [](const BlasReg& a, const BlasReg& b) { return a.hdr_slot < b.hdr_slot; });
  CUDA_OK(c, cudaMalloc(&c->d_spans, c->spans.size() * sizeof(Span)));
  CUDA_OK(c, cudaMemcpyAsync(c->d_spans, c->spans.data(), c->spans.size() * sizeof(Span), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaMalloc(&c->d_blas, std::max<size_t>(c->blas.size(), 1) * sizeof(BlasReg)));
  if (!c->blas.empty()) CUDA_OK(c, cudaMemcpyAsync(c->d_blas, c->blas.data(), c->blas.size() * sizeof(BlasReg), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  c->committed = true;
  return VSRT_OK;
}

int vsrt_form_treelets(vsrt_context* c, const void* tlas, uint32_t max_bytes) {
  if (!c || !tlas) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  if (max_bytes) c->cfg.max_treelet_size = max_bytes;
  return do_form(c, (uint64_t)(uintptr_t)tlas, c->cfg.max_treelet_size);
}

int vsrt_set_treelet_layout_base(vsrt_context* c, uint64_t base) {
  if (!c) return VSRT_E_INVALID;
  if (!c->layout_base_set || c->layout_base != base) c->remap_valid = false;
  c->layout_base = base; c->layout_base_set = true;
  return VSRT_OK;
}

int vsrt_treelet_info_get(vsrt_context* c, vsrt_treelet_info* out) {
  if (!c || !out) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  out->n_treelets = c->fr.n_treelets; out->n_list_entries = c->fr.n_entries; out->n_mapped_nodes = c->fr.n_mapped; out->total_bvh_size = c->fr.total_bvh; out->form_ms = c->fr.ms; out->scratch_bytes = c->fr.peak_scratch_bytes;
  return VSRT_OK;
}

int vsrt_treelet_table(vsrt_context* c, uint64_t* roots, uint64_t* list_offsets, uint64_t* node_addr, uint32_t* node_size) {
  if (!c) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  const int64_t d = formed_delta(c);
  const size_t nt = c->fr.n_treelets;
  if (roots) for (size_t t = 0; t < nt; t++) roots[t] = slot_to_host_h(c, c->h_tl_root[t]) + (uint64_t)d;
  if (list_offsets) memcpy(list_offsets, c->h_tl_off.data(), (nt + 1) * 8);
  for (size_t k = 0; k < c->h_tl_node.size(); k++) {
    if (node_addr) node_addr[k] = entry_dev_addr(c, c->h_tl_node[k], d);
    if (node_size) node_size[k] = ((uint32_t)(c->h_tl_node[k] >> 32) == K_INSTANCE) ? 128u : 64u;
  }
  return VSRT_OK;
}

int vsrt_node_map(vsrt_context* c, uint64_t* node_addr, uint64_t* root_addr) {
  if (!c) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  const int64_t d = formed_delta(c);
  // keys: device address of every list entry; emitted in ascending key order
  std::vector<std::pair<uint64_t, uint64_t>> kv; kv.reserve(c->fr.n_mapped);
  std::vector<uint8_t> seen(c->h_node_tid.size(), 0);
  for (uint64_t e : c->h_tl_node) {
    const uint32_t slot = (uint32_t)e; if (seen[slot]) continue; seen[slot] = 1;
    kv.push_back({ entry_dev_addr(c, e, d), slot_to_host_h(c, c->h_tl_root[c->h_node_tid[slot]]) + (uint64_t)d });
  }
  std::sort(kv.begin(), kv.end());
  for (size_t i = 0; i < kv.size(); i++) { if (node_addr) node_addr[i] = kv[i].first; if (root_addr) root_addr[i] = kv[i].second; }
  return VSRT_OK;
}

int vsrt_treelet_remap(vsrt_context* c, uint64_t base, uint64_t* n_out, uint64_t* orig, uint64_t* mapped) {
  if (!c || !n_out) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  // remapBVHToTreeletLayout (:1473-1509): treelet i at base + i*(max + stride), root first, then the list in order;
  // an entry that is already mapped keeps its first mapping but still advances the cursor (:1501-1503).
  const int64_t d = formed_delta(c);
  const uint64_t pitch = (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride;
  std::vector<std::pair<uint64_t, uint64_t>> kv;
  std::vector<uint8_t> seen(c->h_node_tid.size(), 0);
  for (size_t t = 0; t < c->fr.n_treelets; t++) {
    const uint32_t rslot = c->h_tl_root[t]; const uint64_t rnew = base + t * pitch;
    uint32_t rsize = 64;
    for (uint64_t k = c->h_tl_off[t]; k < c->h_tl_off[t + 1]; k++) if ((uint32_t)c->h_tl_node[k] == rslot && (uint32_t)(c->h_tl_node[k] >> 32) == K_INSTANCE) rsize = 128;
    if (!seen[rslot]) { seen[rslot] = 1; kv.push_back({ slot_to_host_h(c, rslot) + (uint64_t)d, rnew }); }
    uint64_t cur = rnew + rsize;
    for (uint64_t k = c->h_tl_off[t]; k < c->h_tl_off[t + 1]; k++) {
      const uint64_t e = c->h_tl_node[k]; const uint32_t slot = (uint32_t)e;
      if (slot == rslot) continue;
      if (!seen[slot]) { seen[slot] = 1; kv.push_back({ entry_dev_addr(c, e, d), cur }); }
      cur += ((uint32_t)(e >> 32) == K_INSTANCE) ? 128u : 64u;
    }
  }
  std::sort(kv.begin(), kv.end());
  *n_out = kv.size();
  for (size_t i = 0; i < kv.size(); i++) { if (orig) orig[i] = kv[i].first; if (mapped) mapped[i] = kv[i].second; }
  return VSRT_OK;
}

int vsrt_addr_to_treelet(vsrt_context* c, uint64_t addr, uint64_t* root) {
  if (!c || !root) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  uint32_t slot = 0;
  if (!dev_addr_to_slot(c, addr, &slot) || c->h_node_tid[slot] == VSRT_NO_TID) return fail(c, VSRT_E_INVALID, "address 0x%llx is not a BVH node (reference: assert, vulkan_ray_tracing.cc:470)", (unsigned long long)addr);
  *root = slot_to_host_h(c, c->h_tl_root[c->h_node_tid[slot]]) + (uint64_t)formed_delta(c);
  return VSRT_OK;
}

int vsrt_is_treelet_root(vsrt_context* c, uint64_t addr) {
  if (!c) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  uint32_t slot = 0;
  if (!host_to_slot_h(c, addr - (uint64_t)formed_delta(c), &slot)) return 0;
  return std::binary_search(c->h_tl_root.begin(), c->h_tl_root.end(), slot) ? 1 : 0;
}

int vsrt_treelet_metadata_idx(vsrt_context* c, uint64_t root, uint32_t* idx) {
  if (!c || !idx) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  int rc = ensure_mirrors(c); if (rc) return rc;
  uint32_t slot = 0;
  if (!host_to_slot_h(c, root - (uint64_t)formed_delta(c), &slot)) return fail(c, VSRT_E_INVALID, "not a treelet root");
  auto it = std::lower_bound(c->h_tl_root.begin(), c->h_tl_root.end(), slot);
  if (it == c->h_tl_root.end() || *it != slot) return fail(c, VSRT_E_INVALID, "not a treelet root");
  *idx = (uint32_t)(it - c->h_tl_root.begin());
  return VSRT_OK;
}

int vsrt_trace_rays_device(vsrt_context* c, const void* tlas, int mode, uint64_t n, const void* rays_dev, void* stream, uint64_t* n_txn) {
  if (!c || !tlas || (n && !rays_dev)) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  int rc = run_batch(c, (uint64_t)(uintptr_t)tlas, mode, (const vsrt_ray*)rays_dev, n, stream ? (cudaStream_t)stream : c->stream);
  if (n_txn) *n_txn = c->last.n_txn;
  return rc;
}

int vsrt_trace_device_results(vsrt_context* c, vsrt_device_results* out) {
  if (!c || !out) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  int rc = ensure_full_records(c); if (rc) return rc;
  *out = c->last; return VSRT_OK;
}

int vsrt_trace_fetch(vsrt_context* c, vsrt_txn* txns, uint64_t cap, uint64_t* treelet_ids) {
  if (!c) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  { int rc0 = ensure_full_records(c); if (rc0) return rc0; }
  const uint64_t total = c->last.n_txn, m = std::min(total, cap);
  if (txns && m) CUDA_OK(c, cudaMemcpyAsync(txns, c->last.txns, m * sizeof(vsrt_txn), cudaMemcpyDeviceToHost, c->stream));
  if (treelet_ids && m) {
    ArenaView av; int rc = make_view(c, c->last_tlas, &av); if (rc) return rc;
    CUDA_OK(c, c->d_tid_addr.ensure(m));
    const uint64_t pitch = c->cfg.remap_to_treelet_layout ? (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride : 0;
    rc = vsrt_launch_tid_to_addr(av, treelet_view(c), (const uint32_t*)c->last.treelet_ids, m, c->d_tid_addr.p, c->layout_base, pitch, c->stream); if (rc) return fail(c, rc, "tid_to_addr launch failed");
    CUDA_OK(c, cudaMemcpyAsync(treelet_ids, c->d_tid_addr.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return total > cap ? VSRT_E_CAPACITY : VSRT_OK;
}

int vsrt_trace_rays(vsrt_context* c, const void* tlas, int mode, uint64_t n, const vsrt_ray* rays, vsrt_hit* hits, uint64_t* trace_offsets,
                    vsrt_txn* txns, uint64_t txn_capacity, uint64_t* treelet_ids, uint64_t* n_txn) {
  if (!c || !tlas || (n && !rays)) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  {
    // frame-sized batches: traced in windows with the copies overlapped, treelet ids derived on the host (see trace_frame_pipelined)
    uint64_t chunk = 524288;
    if (const char* e = getenv("VSRT_PIPELINE_CHUNK")) chunk = (uint64_t)atoll(e);
    if (chunk && txns && n >= 2 * chunk) {
      const int rcp = trace_frame_pipelined(c, (uint64_t)(uintptr_t)tlas, mode, n, rays, hits, trace_offsets, false, txns, txn_capacity, treelet_ids, n_txn, chunk);
      if (rcp != VSRT_E_UNSUPPORTED) return rcp;
    }
  }
  CUDA_OK(c, c->d_rays.ensure(std::max<uint64_t>(n, 1)));
  uint8_t* const bounce = c->h_pin + vsrt_context::PIN_HEAD; const size_t bounce_bytes = vsrt_context::PIN_BYTES - vsrt_context::PIN_HEAD;
  const bool small = n && n * sizeof(vsrt_ray) <= bounce_bytes && n <= 4096;
  if (small) { memcpy(bounce, rays, n * sizeof(vsrt_ray)); CUDA_OK(c, cudaMemcpyAsync(c->d_rays.p, bounce, n * sizeof(vsrt_ray), cudaMemcpyHostToDevice, c->stream)); }
  else if (n) CUDA_OK(c, cudaMemcpyAsync(c->d_rays.p, rays, n * sizeof(vsrt_ray), cudaMemcpyHostToDevice, c->stream));
  int rc = run_batch(c, (uint64_t)(uintptr_t)tlas, mode, c->d_rays.p, n, c->stream);
  if (n_txn) *n_txn = c->last.n_txn;
  if (rc) return rc;
  {
    // small batch (the 32-lane call): every output through the pinned bounce buffer, one synchronisation
    const uint64_t total = c->last.n_txn, m = std::min(total, txn_capacity);
    const size_t b_hits = hits ? n * sizeof(vsrt_hit) : 0, b_off = trace_offsets ? (n + 1) * 8 : 0, b_txn = txns ? m * sizeof(vsrt_txn) : 0, b_tid = treelet_ids ? m * 8 : 0;
    if (small && b_hits + b_off + b_txn + b_tid <= bounce_bytes) {
      uint8_t* p = bounce;
      if (b_hits) CUDA_OK(c, cudaMemcpyAsync(p, c->d_hits.p, b_hits, cudaMemcpyDeviceToHost, c->stream));
      uint8_t* p_off = p + b_hits;
      if (b_off) CUDA_OK(c, cudaMemcpyAsync(p_off, c->d_offsets.p, b_off, cudaMemcpyDeviceToHost, c->stream));
      uint8_t* p_txn = p_off + b_off;
      if (b_txn) CUDA_OK(c, cudaMemcpyAsync(p_txn, c->last.txns, b_txn, cudaMemcpyDeviceToHost, c->stream));
      uint8_t* p_tid = p_txn + b_txn;
      if (b_tid) {
        ArenaView av; rc = make_view(c, c->last_tlas, &av); if (rc) return rc;
        CUDA_OK(c, c->d_tid_addr.ensure(m));
        const uint64_t pitch = c->cfg.remap_to_treelet_layout ? (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride : 0;
        rc = vsrt_launch_tid_to_addr(av, treelet_view(c), (const uint32_t*)c->last.treelet_ids, m, c->d_tid_addr.p, c->layout_base, pitch, c->stream); if (rc) return fail(c, rc, "tid_to_addr launch failed");
        CUDA_OK(c, cudaMemcpyAsync(p_tid, c->d_tid_addr.p, b_tid, cudaMemcpyDeviceToHost, c->stream));
      }
      CUDA_OK(c, cudaStreamSynchronize(c->stream));
      if (b_hits) memcpy(hits, p, b_hits);
      if (b_off) memcpy(trace_offsets, p_off, b_off);
      if (b_txn) memcpy(txns, p_txn, b_txn);
      if (b_tid) memcpy(treelet_ids, p_tid, b_tid);
      return ((txns || treelet_ids) && total > txn_capacity) ? VSRT_E_CAPACITY : VSRT_OK;
    }
  }
  if (hits && n) CUDA_OK(c, cudaMemcpyAsync(hits, c->d_hits.p, n * sizeof(vsrt_hit), cudaMemcpyDeviceToHost, c->stream));
  if (trace_offsets) CUDA_OK(c, cudaMemcpyAsync(trace_offsets, c->d_offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (txns || treelet_ids) return vsrt_trace_fetch(c, txns, txn_capacity, treelet_ids);
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VSRT_OK;
}

int vsrt_packed_layout_get(vsrt_context* c, const void* tlas, vsrt_packed_layout* out) {
  if (!c || !out) return VSRT_E_INVALID;
  ArenaView av; int rc = make_view(c, (uint64_t)(uintptr_t)tlas, &av); if (rc) return rc;
  if (!av.uniform_delta) return fail(c, VSRT_E_UNSUPPORTED, "packed traces need every BLAS registered with the TLAS's host->device offset");
  if (c->cfg.remap_to_treelet_layout) return fail(c, VSRT_E_UNSUPPORTED, "packed traces carry original addresses; remap_to_treelet_layout is on");
  if (c->spans.size() > 8 || c->spans.empty()) return fail(c, VSRT_E_UNSUPPORTED, "packed traces support 1..8 disjoint host spans (%zu registered)", c->spans.size());
  memset(out, 0, sizeof(*out));
  out->device_delta = av.tlas_delta; out->n_spans = (uint32_t)c->spans.size();
  for (size_t i = 0; i < c->spans.size(); i++) { out->spans[i].host = c->spans[i].host; out->spans[i].slot0 = c->spans[i].slot0; out->spans[i].n_slots = c->spans[i].n_slots; }
  return VSRT_OK;
}

int vsrt_trace_fetch_packed(vsrt_context* c, uint32_t* records, uint64_t cap, uint32_t* treelet_index) {
  if (!c) return VSRT_E_INVALID;
  if (!c->last.trace_offsets) return fail(c, VSRT_E_INVALID, "no trace: call vsrt_trace_rays / vsrt_trace_rays_device first");
  cudaSetDevice(c->device);
  vsrt_packed_layout lay; int rc = vsrt_packed_layout_get(c, (const void*)(uintptr_t)c->last_tlas, &lay); if (rc) return rc;
  const uint64_t total = c->last.n_txn, m = std::min(total, cap);
  if (records && m) {
    if (!c->last_packed_only) {      // (a batch traced in packed form has its packed records in d_packed already: K3 wrote them)
      CUDA_OK(c, c->d_packed.ensure(m));
      rc = vsrt_launch_pack_trace(c->d_stage.p, c->stage_cap, c->d_offsets.p, c->last.n_rays, c->d_packed.p, m, c->stream); if (rc) return fail(c, rc, "pack kernel launch failed");
    }
    CUDA_OK(c, cudaMemcpyAsync(records, c->d_packed.p, m * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if (treelet_index && m) {
    rc = ensure_full_records(c); if (rc) return rc;
    CUDA_OK(c, cudaMemcpyAsync(treelet_index, c->d_tids.p, m * 4, cudaMemcpyDeviceToHost, c->stream));   // traversal order, whatever vsrt_sort_trace did since
  }
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return total > cap ? VSRT_E_CAPACITY : VSRT_OK;
}

}  // extern "C"

namespace {

// The host-buffer calls for a frame-sized batch: the rays are traced in windows of `chunk` rays, and while window k + 1 is uploaded
// and traced the hits / offsets / records of window k go back over PCIe on a second stream.  A frame's trace is hundreds of MB
// even in packed form, so the device->host copy is what bounds a host-side caller; everything else hides beneath it.
//   packed form   records = 4-byte packed records
//   full form     txns = 16-byte records; the 64-bit treelet ids -- a pure function of a record's address -- are not copied
//                 (8 of every 24 bytes) but filled in on the host from a slot -> root table, by worker threads that follow
//                 the copy front (layouts with one span, one host->device offset and no remap; otherwise they are copied)
// Ray ids follow the batch order (windows are traced in order on one stream); every window writes its slice of the frame-sized
// device buffers, so afterwards the context holds the frame as a single batch would have left it.
struct FrameJob {   // one window's host-side work: treelet ids derived from the copied records, or records and ids expanded from packed records
  cudaEvent_t copied; const vsrt_txn* txns; uint64_t* ids; uint64_t n; const uint32_t* packed;
};
void derive_ids(const std::vector<uint64_t>& root_of_slot, uint64_t addr0, const vsrt_txn* txns, uint64_t* ids, uint64_t lo, uint64_t hi) {
  const uint64_t ns = root_of_slot.size();
  for (uint64_t j = lo; j < hi; j++) {
    const uint64_t slot = (txns[j].address - addr0) >> 6;
    const uint64_t v = slot < ns ? root_of_slot[slot] : ~0ull;
#if defined(__x86_64__)
    _mm_stream_si64(reinterpret_cast<long long*>(ids + j), (long long)v);      // written once, never read here: no read-for-ownership
#else
    ids[j] = v;
#endif
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// Full records written on the host from the 4-byte packed records of a window: {address, size, type} is arithmetic on
// (slot, code) for layouts with one span, one host->device offset and no remap, and the 64-bit treelet id a table lookup by
// slot.  The link then carries 4 bytes per record instead of 16 (or 24); the 24 bytes per record the caller asked for are
// produced by worker threads with streaming stores, behind the copy front.
void expand_records(const std::vector<uint64_t>& root_of_slot, uint64_t addr0, const uint32_t* packed, vsrt_txn* txns, uint64_t* ids, uint64_t lo, uint64_t hi) {
  const uint64_t ns = root_of_slot.size();
  static const uint32_t size_of[8] = { 64, 64, 128, 8, 64, 64, 64, 64 };                 // by compact code (vsrt_internal.h): instance leaf 128, descriptor 8
  static const uint32_t type_of[8] = { 0, 1, 2, 3, 4, 5, 6, 1 };                        // code 7 = TLAS internal node -> BVH_INTERNAL_NODE
#if defined(__x86_64__)
  const bool aligned = (reinterpret_cast<uintptr_t>(txns) & 15u) == 0;
#endif
  for (uint64_t j = lo; j < hi; j++) {
    const uint32_t rec = packed[j], code = rec & 7u; const uint64_t slot = rec >> 3;
    const uint64_t address = addr0 + slot * 64u;
    if (txns) {
#if defined(__x86_64__)
      if (aligned) _mm_stream_si128(reinterpret_cast<__m128i*>(txns + j), _mm_set_epi32((int)type_of[code], (int)size_of[code], (int)(address >> 32), (int)(uint32_t)address));
      else
#endif
      { txns[j].address = address; txns[j].size = size_of[code]; txns[j].type = type_of[code]; }
    }
    if (ids) {
      const uint64_t v = slot < ns ? root_of_slot[slot] : ~0ull;
#if defined(__x86_64__)
      _mm_stream_si64(reinterpret_cast<long long*>(ids + j), (long long)v);
#else
      ids[j] = v;
#endif
    }
  }
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

int trace_frame_pipelined(vsrt_context* c, uint64_t tlas, int mode, uint64_t n, const vsrt_ray* rays, vsrt_hit* hits, uint64_t* trace_offsets,
                          bool packed, void* records_out, uint64_t capacity, uint64_t* treelet_ids, uint64_t* n_txn, uint64_t chunk) {
  if (!c->copy_stream) {
    CUDA_OK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)); CUDA_OK(c, cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming)); CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming)); }
    CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  }
  CUDA_OK(c, c->d_rays.ensure(n));
  int rc = do_form(c, tlas, c->cfg.max_treelet_size); if (rc) return rc;
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  // host-derived treelet ids: slot -> root address table (once per formation), addresses start at the span's first byte
  // full form, simple layout: the windows run in packed form and the host expands them (VSRT_HOST_EXPAND=0: records copied as
  // 16-byte records, ids derived from their addresses -- the first round-2 scheme, kept for A/B)
  bool host_expand = !packed && records_out && c->spans.size() == 1 && !c->cfg.remap_to_treelet_layout;
  if (const char* e = getenv("VSRT_HOST_EXPAND")) host_expand = host_expand && atoi(e) != 0;
  const bool derive = !packed && (treelet_ids || host_expand) && c->spans.size() == 1 && !c->cfg.remap_to_treelet_layout;
  uint64_t addr0 = 0; unsigned threads = 1;
  if (derive) {
    ArenaView av; rc = make_view(c, tlas, &av); if (rc) return rc;
    if (!av.uniform_delta) return VSRT_E_UNSUPPORTED;       // (the caller falls back to the one-batch path)
    rc = ensure_mirrors(c); if (rc) return rc;
    if (c->h_root_of_slot.size() != c->h_node_tid.size()) {
      c->h_root_of_slot.resize(c->h_node_tid.size());
      const int64_t d = formed_delta(c);
      for (size_t sl = 0; sl < c->h_node_tid.size(); sl++) {
        const uint32_t t = c->h_node_tid[sl];
        c->h_root_of_slot[sl] = t == VSRT_NO_TID ? ~0ull : slot_to_host_h(c, c->h_tl_root[t]) + (uint64_t)d;
      }
    }
    addr0 = c->spans[0].host + (uint64_t)av.tlas_delta;
    // three quarters of the cores: the thread that launches the next window and the copy engine's completion path need the rest
    // (16-core box, e2e M rays/s at 4 / 8 / 12 / 15 / 16 / 20 workers: 42 / 59 / 92 / 86 / 88 / 75)
    { const unsigned hc = std::max(1u, std::thread::hardware_concurrency()); threads = std::max(1u, std::min(hc - hc / 4, 48u)); }
    if (const char* e = getenv("VSRT_HOST_THREADS")) threads = (unsigned)std::max(1, atoi(e));
  }
  // a stage-capacity change in a later window invalidates the earlier windows' staging: the frame is restarted then, from a
  // copy of the counters and histograms taken here (cheap: a few MB device-to-device)
  const size_t bak_bytes = sizeof(DevCounters) + (size_t)c->hist_n * 8 + (c->node_hist_on ? (size_t)c->node_hist_n * 8 : 0);
  CUDA_OK(c, c->d_frame_bak.ensure(bak_bytes));
  CUDA_OK(c, cudaMemcpyAsync(c->d_frame_bak.p, c->d_counters, sizeof(DevCounters), cudaMemcpyDeviceToDevice, c->stream));
  if (c->hist_n) CUDA_OK(c, cudaMemcpyAsync(c->d_frame_bak.p + sizeof(DevCounters), c->d_hist.p, (size_t)c->hist_n * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (c->node_hist_on && c->node_hist_n) CUDA_OK(c, cudaMemcpyAsync(c->d_frame_bak.p + sizeof(DevCounters) + (size_t)c->hist_n * 8, c->d_node_hist.p, (size_t)c->node_hist_n * 8, cudaMemcpyDeviceToDevice, c->stream));
  const DevCounters h_prev_bak = c->h_prev;
  std::vector<cudaEvent_t> win_ev;
  std::vector<FrameJob> jobs;
  // a window's host-side work goes to the context's worker pool and this thread carries on with the next window; the workers
  // wait for the window's copy themselves.  One job at a time: submitting waits for the previous window's job, which is also
  // what makes a staging buffer free again by the time the window after next is copied into it.
  if (derive && (!c->pool || c->pool->th.size() != threads)) { delete c->pool; c->pool = new HostPool(threads); }
  const std::vector<uint64_t>* const rtab = &c->h_root_of_slot; const int device = c->device;
  auto run_job = [&](const FrameJob& j) {
    // the window's records are handed out in pieces of 32,768 from a shared cursor rather than as one equal slice per worker: a
    // worker whose core is also serving something else (the launching thread, the driver's completion path, another rank) then
    // takes fewer pieces instead of making the whole window wait for its slice (e2e of one box, one run after the other, with
    // equal slices: 65 .. 95 M rays/s)
    std::atomic<uint64_t>* const cursor = &c->pool->next;
    c->pool->submit([j, rtab, addr0, device, cursor](unsigned, unsigned) {
      cudaSetDevice(device);
      cudaEventSynchronize(j.copied);
      constexpr uint64_t PIECE = 32768;
      for (;;) {
        const uint64_t lo = cursor->fetch_add(PIECE, std::memory_order_relaxed);
        if (lo >= j.n) break;
        const uint64_t hi = std::min(lo + PIECE, j.n);
        if (j.packed) expand_records(*rtab, addr0, j.packed, const_cast<vsrt_txn*>(j.txns), j.ids, lo, hi);
        else if (j.ids) derive_ids(*rtab, addr0, j.txns, j.ids, lo, hi);
      }
    });
  };
  const uint64_t rec_size = packed ? 4 : sizeof(vsrt_txn);
  for (int attempt = 0; attempt < 12; attempt++) {
    uint64_t base = 0; bool overflow = false, restart = false;
    jobs.clear();
    CUDA_OK(c, cudaMemcpyAsync(c->d_rays.p, rays, std::min(chunk, n) * sizeof(vsrt_ray), cudaMemcpyHostToDevice, c->up_stream));
    CUDA_OK(c, cudaEventRecord(c->ev_up[0], c->up_stream));
    for (uint64_t k = 0; k < n_chunks; k++) {
      const uint64_t r0 = k * chunk, m = std::min(chunk, n - r0); const int b = (int)(k & 1);
      CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->ev_up[b], 0));
      if (k + 1 < n_chunks) {   // next window's rays go up while this one is traced
        const uint64_t r1 = r0 + chunk, m1 = std::min(chunk, n - r1);
        CUDA_OK(c, cudaMemcpyAsync(c->d_rays.p + r1, rays + r1, m1 * sizeof(vsrt_ray), cudaMemcpyHostToDevice, c->up_stream));
        CUDA_OK(c, cudaEventRecord(c->ev_up[b ^ 1], c->up_stream));
      }
      const uint32_t cap_before = c->stage_cap;
      rc = run_batch(c, tlas, mode, c->d_rays.p + r0, m, c->stream, packed || host_expand, r0, n, base);      // ends with a host synchronisation of c->stream
      if (rc) break;
      if (c->stage_cap != cap_before && k > 0) { restart = true; break; }
      const uint64_t total = c->last.n_txn - base;
      CUDA_OK(c, cudaEventRecord(c->ev_ready, c->stream));
      CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_ready, 0));
      if (hits) CUDA_OK(c, cudaMemcpyAsync(hits + r0, c->d_hits.p + r0, m * sizeof(vsrt_hit), cudaMemcpyDeviceToHost, c->copy_stream));
      // (the window's last offset is the next window's first and is rewritten by it: only the last window copies it)
      if (trace_offsets) CUDA_OK(c, cudaMemcpyAsync(trace_offsets + r0, c->d_offsets.p + r0, (m + (k + 1 == n_chunks ? 1 : 0)) * 8, cudaMemcpyDeviceToHost, c->copy_stream));
      uint64_t mrec = 0;
      if (records_out && base < capacity) {
        mrec = std::min(total, capacity - base);
        if (host_expand) {
          // the window's packed records land in one of two pinned staging buffers of the library; the workers expand from there
          HostStage& hs = c->h_stage[b];
          if (hs.cap < mrec) {
            if (hs.p) cudaFreeHost(hs.p);
            hs.p = nullptr; hs.cap = 0;
            const uint64_t want = mrec + mrec / 4 + 4096;
            CUDA_OK(c, cudaMallocHost(&hs.p, want * 4)); hs.cap = want;
          }
          if (mrec) CUDA_OK(c, cudaMemcpyAsync(hs.p, c->d_packed.p + base, mrec * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        } else {
          const void* src = packed ? (const void*)(c->d_packed.p + base) : (const void*)(c->d_txns.p + base);
          if (mrec) CUDA_OK(c, cudaMemcpyAsync((uint8_t*)records_out + base * rec_size, src, mrec * rec_size, cudaMemcpyDeviceToHost, c->copy_stream));
        }
      }
      if (base + total > capacity) overflow = true;
      if (win_ev.size() <= k) { cudaEvent_t e; CUDA_OK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); win_ev.push_back(e); }
      CUDA_OK(c, cudaEventRecord(win_ev[k], c->copy_stream));
      if (derive && mrec) {
        // the PREVIOUS window is expanded (or its ids derived) now, while this window's copy is in flight: the workers need its records in host memory
        jobs.push_back(FrameJob{ win_ev[k], (const vsrt_txn*)records_out + base, treelet_ids ? treelet_ids + base : nullptr, mrec, host_expand ? c->h_stage[b].p : nullptr });
        run_job(jobs.back());
      }
      base += total;
    }
    if (c->pool) c->pool->wait();
    cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->up_stream);
    if (restart) {
      CUDA_OK(c, cudaMemcpyAsync(c->d_counters, c->d_frame_bak.p, sizeof(DevCounters), cudaMemcpyDeviceToDevice, c->stream));
      if (c->hist_n) CUDA_OK(c, cudaMemcpyAsync(c->d_hist.p, c->d_frame_bak.p + sizeof(DevCounters), (size_t)c->hist_n * 8, cudaMemcpyDeviceToDevice, c->stream));
      if (c->node_hist_on && c->node_hist_n) CUDA_OK(c, cudaMemcpyAsync(c->d_node_hist.p, c->d_frame_bak.p + sizeof(DevCounters) + (size_t)c->hist_n * 8, (size_t)c->node_hist_n * 8, cudaMemcpyDeviceToDevice, c->stream));
      CUDA_OK(c, cudaStreamSynchronize(c->stream));
      c->h_prev = h_prev_bak;
      continue;
    }
    for (cudaEvent_t e : win_ev) cudaEventDestroy(e);
    if (n_txn) *n_txn = rc ? 0 : base;
    if (rc) return rc;
    if (!packed && treelet_ids && !derive) {
      // layouts the host table does not cover: ids expanded on the device and copied (8 more bytes per record over the link)
      const uint64_t m = std::min(base, capacity);
      if (m) {
        ArenaView av; rc = make_view(c, c->last_tlas, &av); if (rc) return rc;
        CUDA_OK(c, c->d_tid_addr.ensure(m));
        const uint64_t pitch = c->cfg.remap_to_treelet_layout ? (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride : 0;
        rc = vsrt_launch_tid_to_addr(av, treelet_view(c), (const uint32_t*)c->last.treelet_ids, m, c->d_tid_addr.p, c->layout_base, pitch, c->stream); if (rc) return fail(c, rc, "tid_to_addr launch failed");
        CUDA_OK(c, cudaMemcpyAsync(treelet_ids, c->d_tid_addr.p, m * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
      }
    }
    return (records_out && overflow) ? VSRT_E_CAPACITY : VSRT_OK;
  }
  for (cudaEvent_t e : win_ev) cudaEventDestroy(e);
  return fail(c, VSRT_E_CAPACITY, "per-ray trace staging kept growing");
}
}  // namespace

extern "C" {

int vsrt_trace_rays_packed(vsrt_context* c, const void* tlas, int mode, uint64_t n, const vsrt_ray* rays, vsrt_hit* hits, uint64_t* trace_offsets,
                           uint32_t* records, uint64_t capacity, uint32_t* treelet_index, uint64_t* n_txn) {
  if (!c || !tlas || (n && !rays)) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  // frame-sized batches whose records are wanted without the treelet-index stream (derive it from vsrt_node_treelet_table) are
  // traced in chunks with the copies overlapped; VSRT_PIPELINE_CHUNK sets the chunk size in rays (0 = never)
  uint64_t chunk = 524288;
  if (const char* e = getenv("VSRT_PIPELINE_CHUNK")) chunk = (uint64_t)atoll(e);
  if (chunk && records && !treelet_index && n >= 2 * chunk) {
    vsrt_packed_layout lay; int rc0 = vsrt_packed_layout_get(c, tlas, &lay); if (rc0) return rc0;
    return trace_frame_pipelined(c, (uint64_t)(uintptr_t)tlas, mode, n, rays, hits, trace_offsets, true, records, capacity, nullptr, n_txn, chunk);
  }
  CUDA_OK(c, c->d_rays.ensure(std::max<uint64_t>(n, 1)));
  if (n) CUDA_OK(c, cudaMemcpyAsync(c->d_rays.p, rays, n * sizeof(vsrt_ray), cudaMemcpyHostToDevice, c->stream));
  const bool packed_only = !treelet_index;
  int rc = run_batch(c, (uint64_t)(uintptr_t)tlas, mode, c->d_rays.p, n, c->stream, packed_only);
  if (n_txn) *n_txn = c->last.n_txn;
  if (rc) return rc;
  if (hits && n) CUDA_OK(c, cudaMemcpyAsync(hits, c->d_hits.p, n * sizeof(vsrt_hit), cudaMemcpyDeviceToHost, c->stream));
  if (trace_offsets) CUDA_OK(c, cudaMemcpyAsync(trace_offsets, c->d_offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (records || treelet_index) return vsrt_trace_fetch_packed(c, records, capacity, treelet_index);
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return VSRT_OK;
}

int vsrt_node_treelet_table(vsrt_context* c, uint32_t* treelet_of_slot, uint64_t capacity, uint64_t* n_slots) {
  if (!c) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  cudaSetDevice(c->device);
  int rc = ensure_mirrors(c); if (rc) return rc;
  if (n_slots) *n_slots = c->h_node_tid.size();
  if (!treelet_of_slot) return VSRT_OK;
  if (capacity < c->h_node_tid.size()) return fail(c, VSRT_E_CAPACITY, "the table has %zu entries", c->h_node_tid.size());
  memcpy(treelet_of_slot, c->h_node_tid.data(), c->h_node_tid.size() * 4);
  return VSRT_OK;
}

void vsrt_unpack_txns(const vsrt_packed_layout* layout, const uint32_t* records, uint64_t n, vsrt_txn* out) {
  for (uint64_t i = 0; i < n; i++) vsrt_unpack_txn(layout, records[i], &out[i]);
}

int vsrt_trace_ray_warp(vsrt_context* c, const void* tlas, uint32_t active_mask, const vsrt_ray rays[32], vsrt_hit hits[32],
                        uint32_t txn_counts[32], vsrt_txn* txns, uint64_t txn_capacity, uint64_t* n_txn) {
  if (!c || !tlas || !rays) return VSRT_E_INVALID;
  // lanes are processed in lane order, active lanes only (core_t::execute_warp_inst_t, abstract_hardware_model.cc:3052-3063)
  vsrt_ray packed[32]; int lane_of[32]; uint32_t n = 0;
  for (int l = 0; l < 32; l++) if (active_mask & (1u << l)) { packed[n] = rays[l]; lane_of[n] = l; n++; }
  vsrt_hit ph[32]; uint64_t off[33];
  int rc = vsrt_trace_rays(c, tlas, c->cfg.treelet_based_traversal ? VSRT_MODE_TREELET : VSRT_MODE_DFS, n, packed, ph, off, txns, txn_capacity, nullptr, n_txn);
  if (rc && rc != VSRT_E_CAPACITY) return rc;
  if (txn_counts) memset(txn_counts, 0, 32 * sizeof(uint32_t));
  for (uint32_t i = 0; i < n; i++) { if (hits) hits[lane_of[i]] = ph[i]; if (txn_counts) txn_counts[lane_of[i]] = (uint32_t)(off[i + 1] - off[i]); }
  return rc;
}

int vsrt_get_counters(vsrt_context* c, vsrt_counters* out) {
  if (!c || !out) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  static_assert(sizeof(vsrt_counters) == sizeof(DevCounters), "counter layouts must match");
  CUDA_OK(c, cudaMemcpy(out, c->d_counters, sizeof(DevCounters), cudaMemcpyDeviceToHost));
  return VSRT_OK;
}
int vsrt_reset_counters(vsrt_context* c) {
  if (!c) return VSRT_E_INVALID;
  cudaSetDevice(c->device);
  CUDA_OK(c, cudaMemset(c->d_counters, 0, sizeof(DevCounters)));
  c->h_prev = DevCounters{};
  if (c->hist_n) CUDA_OK(c, cudaMemset(c->d_hist.p, 0, (size_t)c->hist_n * 8));
  if (c->node_hist_n) CUDA_OK(c, cudaMemset(c->d_node_hist.p, 0, (size_t)c->node_hist_n * 8));
  return vsrt_comm_counters_reset(c);
}
int vsrt_get_treelet_histogram(vsrt_context* c, uint64_t* hist, uint64_t capacity) {
  if (!c || !hist) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  if (capacity < c->hist_n) return fail(c, VSRT_E_CAPACITY, "histogram needs %u entries", c->hist_n);
  cudaSetDevice(c->device);
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  CUDA_OK(c, cudaMemcpy(hist, c->d_hist.p, (size_t)c->hist_n * 8, cudaMemcpyDeviceToHost));
  return VSRT_OK;
}
// debug export (not in vsrt.h): statistics of the last batch the treelet-binned kernel traced (TbParams::stats + rounds)
int vsrt_debug_tb_stats(vsrt_context* c, unsigned long long out[8]) {
  if (!c || !out) return VSRT_E_INVALID;
  memcpy(out, c->tb_stats, sizeof(c->tb_stats)); return VSRT_OK;
}
int vsrt_enable_node_histogram(vsrt_context* c, int enable) {
  if (!c) return VSRT_E_INVALID;
  c->node_hist_on = enable != 0;
  return VSRT_OK;
}
int vsrt_get_node_histogram(vsrt_context* c, uint64_t* visits_of_slot, uint64_t capacity, uint64_t* n_slots) {
  if (!c) return VSRT_E_INVALID;
  const uint64_t ns = c->arena_bytes / 64;
  if (n_slots) *n_slots = ns;
  if (!visits_of_slot) return VSRT_OK;
  if (!c->node_hist_on) return fail(c, VSRT_E_INVALID, "the node-visit histogram is off: vsrt_enable_node_histogram(ctx, 1) before tracing");
  if (capacity < ns) return fail(c, VSRT_E_CAPACITY, "the node histogram has %llu entries", (unsigned long long)ns);
  cudaSetDevice(c->device);
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (c->node_hist_n != ns) { memset(visits_of_slot, 0, ns * 8); return VSRT_OK; }      // nothing traced yet
  CUDA_OK(c, cudaMemcpy(visits_of_slot, c->d_node_hist.p, ns * 8, cudaMemcpyDeviceToHost));
  return VSRT_OK;
}
int vsrt_counters_device(vsrt_context* c, void** counters_dev, void** hist_dev, uint64_t* n_treelets) {
  if (!c) return VSRT_E_INVALID;
  if (counters_dev) *counters_dev = c->d_counters;
  if (hist_dev) *hist_dev = c->d_hist.p;
  if (n_treelets) *n_treelets = c->hist_n;
  return VSRT_OK;
}

// ---------------------------------------------------------------- RT-unit replay helpers
int vsrt_sort_trace(vsrt_context* c, int method) {
  if (!c) return VSRT_E_INVALID;
  if (method != 0 && method != 1) return fail(c, VSRT_E_INVALID, "sort method must be 0 (strict) or 1 (loose), -sort_method of gpgpusim.config");
  if (!c->formed || !c->last.trace_offsets) return fail(c, VSRT_E_INVALID, "no trace to sort: call vsrt_trace_rays / vsrt_trace_rays_device first");
  cudaSetDevice(c->device);
  { int rc0 = ensure_full_records(c); if (rc0) return rc0; }
  const uint64_t n = c->last.n_rays, total = c->last.n_txn;
  CUDA_OK(c, c->d_txns_sorted.ensure(std::max<uint64_t>(total, 1))); CUDA_OK(c, c->d_tids_sorted.ensure(std::max<uint64_t>(total, 1)));
  CUDA_OK(c, c->d_sort_keys.ensure(std::max<uint64_t>(total, 1)));
  if (method == 0 && !c->d_inv_off) {
    int rc = vsrt_launch_build_inverse(c->fo, c->fr.n_treelets, c->fr.n_entries, (uint32_t)(c->arena_bytes / 64), &c->d_inv_off, &c->d_inv, c->stream);
    if (rc) return fail(c, rc, "building the inverted treelet lists failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  int rc = vsrt_launch_sort_trace(method, c->d_offsets.p, n, c->d_txns.p, c->d_tids.p, c->d_stage.p, c->stage_cap, c->d_inv_off, c->d_inv,
                                  c->d_txns_sorted.p, c->d_tids_sorted.p, c->d_sort_keys.p, c->stream);
  if (rc) return fail(c, rc, "sort kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  c->last.txns = c->d_txns_sorted.p; c->last.treelet_ids = c->d_tids_sorted.p;
  return VSRT_OK;
}

}  // extern "C"

namespace {
uint64_t treelet_root_address(vsrt_context* c, uint32_t t) {
  if (c->cfg.remap_to_treelet_layout) return c->layout_base + (uint64_t)t * ((uint64_t)c->formed_budget + c->cfg.treelet_remap_stride);
  return slot_to_host_h(c, c->h_tl_root[t]) + (uint64_t)formed_delta(c);
}
bool treelet_index_of(vsrt_context* c, uint64_t root, uint32_t* idx) {
  if (c->cfg.remap_to_treelet_layout) {
    const uint64_t pitch = (uint64_t)c->formed_budget + c->cfg.treelet_remap_stride;
    if (root < c->layout_base || (root - c->layout_base) % pitch) return false;
    const uint64_t t = (root - c->layout_base) / pitch; if (t >= c->fr.n_treelets) return false;
    *idx = (uint32_t)t; return true;
  }
  uint32_t slot = 0;
  if (!host_to_slot_h(c, root - (uint64_t)formed_delta(c), &slot)) return false;
  auto it = std::lower_bound(c->h_tl_root.begin(), c->h_tl_root.end(), slot);
  if (it == c->h_tl_root.end() || *it != slot) return false;
  *idx = (uint32_t)(it - c->h_tl_root.begin()); return true;
}
template <typename T> cudaError_t upload(T** dev, const T* host, size_t n, cudaStream_t st) {
  *dev = nullptr; if (!host || !n) return cudaSuccess;
  cudaError_t e = cudaMalloc(dev, n * sizeof(T)); if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice, st);
}
}  // namespace

extern "C" {

int vsrt_prefetch_vote(vsrt_context* c, const vsrt_prefetch_config* cfg, uint64_t n_groups, const uint64_t* group_offsets,
                       const uint64_t* ray_ids, const uint32_t* front, vsrt_prefetch_decision* decisions) {
  if (!c || !cfg || (n_groups && (!group_offsets || !decisions))) return VSRT_E_INVALID;
  if (cfg->heuristic > 3) return fail(c, VSRT_E_INVALID, "treelet_prefetch_heuristic must be 0..3");
  if (!c->formed || !c->last.trace_offsets) return fail(c, VSRT_E_INVALID, "no trace to vote on: call vsrt_trace_rays / vsrt_trace_rays_device first");
  if (n_groups == 0) return VSRT_OK;
  cudaSetDevice(c->device);
  int rc = ensure_mirrors(c); if (rc) return rc;
  rc = ensure_full_records(c); if (rc) return rc;
  const uint64_t n_ids = group_offsets[n_groups];
  if (!ray_ids && n_ids > c->last.n_rays) return fail(c, VSRT_E_INVALID, "groups cover %llu rays, the last batch has %llu", (unsigned long long)n_ids, (unsigned long long)c->last.n_rays);
  uint64_t* d_go = nullptr; uint64_t* d_ids = nullptr; uint32_t* d_front = nullptr; vsrt_prefetch_decision* d_out = nullptr;
  bool ok = upload(&d_go, group_offsets, n_groups + 1, c->stream) == cudaSuccess && upload(&d_ids, ray_ids, n_ids, c->stream) == cudaSuccess &&
            upload(&d_front, front, c->last.n_rays, c->stream) == cudaSuccess && cudaMalloc(&d_out, n_groups * sizeof(vsrt_prefetch_decision)) == cudaSuccess;
  rc = ok ? vsrt_launch_prefetch_vote((const uint64_t*)c->last.trace_offsets, (const uint32_t*)c->last.treelet_ids, c->last.n_rays, d_go, group_offsets, d_ids, d_front,
                                      n_groups, c->fo, c->fr.n_treelets, cfg->heuristic, cfg->threshold, d_out, c->stream) : VSRT_E_CUDA;
  if (rc == VSRT_OK && (cudaMemcpyAsync(decisions, d_out, n_groups * sizeof(vsrt_prefetch_decision), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                        cudaStreamSynchronize(c->stream) != cudaSuccess)) rc = VSRT_E_CUDA;
  cudaFree(d_go); cudaFree(d_ids); cudaFree(d_front); cudaFree(d_out);
  if (rc) return fail(c, rc, "prefetch vote failed: %s", cudaGetErrorString(cudaGetLastError()));
  for (uint64_t g = 0; g < n_groups; g++) decisions[g].treelet_root = decisions[g].treelet_root == ~0ull ? 0ull : treelet_root_address(c, (uint32_t)decisions[g].treelet_root);
  return VSRT_OK;
}

int vsrt_prefetch_chunks(vsrt_context* c, const vsrt_prefetch_config* cfg, uint64_t n_groups, const vsrt_prefetch_decision* decisions,
                         uint64_t* chunk_offsets, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t capacity, uint64_t* n_chunks) {
  if (!c || !cfg || (n_groups && !decisions)) return VSRT_E_INVALID;
  if (!c->formed) return fail(c, VSRT_E_INVALID, "treelets not formed");
  if (n_chunks) *n_chunks = 0;
  if (n_groups == 0) { if (chunk_offsets) chunk_offsets[0] = 0; return VSRT_OK; }
  cudaSetDevice(c->device);
  int rc = ensure_mirrors(c); if (rc) return rc;
  if (c->cfg.remap_to_treelet_layout && !c->remap_valid) return fail(c, VSRT_E_INVALID, "remapped layout: trace once (or set the layout base) before asking for prefetch chunks");
  std::vector<vsrt_prefetch_decision> dec(decisions, decisions + n_groups);
  for (auto& d : dec) {
    uint32_t idx = 0;
    if (d.treelet_root == 0 || !d.submit) { d.treelet_root = ~0ull; d.submit = 0; continue; }
    if (!treelet_index_of(c, d.treelet_root, &idx)) return fail(c, VSRT_E_INVALID, "decision names 0x%llx, which is not a treelet root", (unsigned long long)d.treelet_root);
    d.treelet_root = idx;
  }
  ArenaView av; rc = make_view(c, c->formed_tlas, &av); if (rc) return rc;
  const uint32_t per_meta = (c->formed_budget / 64u) * 4u;       // per_treelet_metadata_size, vulkan_ray_tracing.cc:1601-1602
  vsrt_prefetch_decision* d_dec = nullptr; uint32_t* d_cnt = nullptr; uint64_t* d_off = nullptr; void* d_tmp = nullptr; uint64_t* d_ca = nullptr; uint64_t* d_co = nullptr;
  uint64_t total = 0;
  bool ok = upload(&d_dec, dec.data(), n_groups, c->stream) == cudaSuccess && cudaMalloc(&d_cnt, n_groups * 4) == cudaSuccess &&
            cudaMalloc(&d_off, (n_groups + 1) * 8) == cudaSuccess && cudaMalloc(&d_tmp, vsrt_scan_tmp_bytes(n_groups)) == cudaSuccess;
  rc = ok ? VSRT_OK : VSRT_E_CUDA;
  const uint64_t* remap = c->cfg.remap_to_treelet_layout ? c->d_remap.p : nullptr;
  if (rc == VSRT_OK) rc = vsrt_launch_prefetch_chunks(false, av, treelet_view(c), c->fo, remap, d_dec, n_groups, cfg->load_treelet_metadata, per_meta, cfg->treelet_metadata_base,
                                                      d_cnt, nullptr, nullptr, nullptr, 0, c->stream);
  if (rc == VSRT_OK) rc = vsrt_launch_scan(d_cnt, n_groups, d_off, d_tmp, c->stream);
  if (rc == VSRT_OK && (cudaMemcpyAsync(&total, d_off + n_groups, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)) rc = VSRT_E_CUDA;
  if (rc == VSRT_OK && chunk_offsets && cudaMemcpy(chunk_offsets, d_off, (n_groups + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = VSRT_E_CUDA;
  if (n_chunks) *n_chunks = total;
  const uint64_t m = std::min(total, capacity);
  if (rc == VSRT_OK && m && chunk_addr && chunk_owner) {
    ok = cudaMalloc(&d_ca, total * 8) == cudaSuccess && cudaMalloc(&d_co, total * 8) == cudaSuccess;
    rc = ok ? vsrt_launch_prefetch_chunks(true, av, treelet_view(c), c->fo, remap, d_dec, n_groups, cfg->load_treelet_metadata, per_meta, cfg->treelet_metadata_base,
                                          d_cnt, d_off, d_ca, d_co, total, c->stream) : VSRT_E_CUDA;
    if (rc == VSRT_OK && (cudaMemcpyAsync(chunk_addr, d_ca, m * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                          cudaMemcpyAsync(chunk_owner, d_co, m * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)) rc = VSRT_E_CUDA;
  }
  cudaFree(d_dec); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_ca); cudaFree(d_co);
  if (rc) return fail(c, rc, "prefetch chunk generation failed: %s", cudaGetErrorString(cudaGetLastError()));
  return total > capacity ? VSRT_E_CAPACITY : VSRT_OK;
}

int vsrt_schedule_pick(vsrt_context* c, int scheduler, uint64_t n_units, const uint64_t* unit_warp_offsets, const uint64_t* warp_ray_ids,
                       const uint8_t* stalled, const uint64_t* last_prefetched, const uint32_t* front, int64_t* pick) {
  if (!c || (n_units && (!unit_warp_offsets || !warp_ray_ids || !pick))) return VSRT_E_INVALID;
  if (scheduler < 0 || scheduler > 2) return fail(c, VSRT_E_INVALID, "treelet_scheduler must be 0, 1 or 2");
  if (!c->formed || !c->last.trace_offsets) return fail(c, VSRT_E_INVALID, "no trace: call vsrt_trace_rays / vsrt_trace_rays_device first");
  if (n_units == 0) return VSRT_OK;
  cudaSetDevice(c->device);
  int rc = ensure_mirrors(c); if (rc) return rc;
  rc = ensure_full_records(c); if (rc) return rc;
  const uint64_t n_warps = unit_warp_offsets[n_units];
  std::vector<uint32_t> target(n_units, VSRT_NO_TID);
  if (last_prefetched) for (uint64_t u = 0; u < n_units; u++) { uint32_t t; if (last_prefetched[u] && treelet_index_of(c, last_prefetched[u], &t)) target[u] = t; }
  uint64_t* d_uo = nullptr; uint64_t* d_ids = nullptr; uint8_t* d_st = nullptr; uint32_t* d_tg = nullptr; uint32_t* d_front = nullptr; int64_t* d_pick = nullptr;
  bool ok = upload(&d_uo, unit_warp_offsets, n_units + 1, c->stream) == cudaSuccess && upload(&d_ids, warp_ray_ids, n_warps * 32, c->stream) == cudaSuccess &&
            upload(&d_st, stalled, n_warps, c->stream) == cudaSuccess && upload(&d_tg, target.data(), n_units, c->stream) == cudaSuccess &&
            upload(&d_front, front, c->last.n_rays, c->stream) == cudaSuccess && cudaMalloc(&d_pick, n_units * 8) == cudaSuccess;
  rc = ok ? vsrt_launch_schedule_pick((const uint64_t*)c->last.trace_offsets, (const uint32_t*)c->last.treelet_ids, c->last.n_rays, d_uo, d_ids, d_st, d_tg, d_front,
                                      n_units, scheduler, d_pick, c->stream) : VSRT_E_CUDA;
  if (rc == VSRT_OK && (cudaMemcpyAsync(pick, d_pick, n_units * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)) rc = VSRT_E_CUDA;
  cudaFree(d_uo); cudaFree(d_ids); cudaFree(d_st); cudaFree(d_tg); cudaFree(d_front); cudaFree(d_pick);
  if (rc) return fail(c, rc, "schedule pick failed: %s", cudaGetErrorString(cudaGetLastError()));
  return VSRT_OK;
}

int vsrt_table_events(vsrt_context* c, const uint8_t* tid_x, uint64_t* event_offsets, vsrt_table_event* events, vsrt_hit* anyhit,
                      uint64_t capacity, uint64_t* n_events) {
  if (!c) return VSRT_E_INVALID;
  if (n_events) *n_events = 0;
  if (!c->formed || !c->last.trace_offsets || !c->last_rays) return fail(c, VSRT_E_INVALID, "no trace: call vsrt_trace_rays / vsrt_trace_rays_device first (the rays must still be resident)");
  cudaSetDevice(c->device);
  const uint64_t n = c->last.n_rays;
  if (n == 0) { if (event_offsets) event_offsets[0] = 0; return VSRT_OK; }
  ArenaView av; int rc = make_view(c, c->last_tlas, &av); if (rc) return rc;
  cudaStream_t st = c->stream;
  uint8_t* d_tid = nullptr; uint32_t* d_cnt = nullptr; uint64_t* d_off = nullptr; void* d_tmp = nullptr; vsrt_table_event* d_ev = nullptr; vsrt_hit* d_ah = nullptr;
  uint64_t total = 0;
  bool ok = upload(&d_tid, tid_x, n, st) == cudaSuccess && cudaMalloc(&d_cnt, n * 4) == cudaSuccess && cudaMalloc(&d_off, (n + 1) * 8) == cudaSuccess &&
            cudaMalloc(&d_tmp, vsrt_scan_tmp_bytes(n)) == cudaSuccess;
  TableParams tp; tp.av = av; tp.rays = c->last_rays; tp.n_rays = n; tp.stage = c->d_stage.p; tp.cap = c->stage_cap; tp.mode = (uint32_t)c->last_mode;
  tp.counts = c->d_counts.p; tp.nproc = c->d_nproc.p; tp.tid_x = d_tid; tp.ev_counts = d_cnt; tp.ev_offsets = d_off; tp.events = nullptr; tp.anyhit = nullptr; tp.capacity = 0;
  rc = ok ? vsrt_launch_table_events(false, tp, nullptr, st) : VSRT_E_CUDA;
  if (rc == VSRT_OK) rc = vsrt_launch_scan(d_cnt, n, d_off, d_tmp, st);
  if (rc == VSRT_OK && (cudaMemcpyAsync(&total, d_off + n, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) rc = VSRT_E_CUDA;
  if (rc == VSRT_OK && event_offsets && cudaMemcpy(event_offsets, d_off, (n + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = VSRT_E_CUDA;
  if (n_events) *n_events = total;
  const uint64_t m = std::min(total, capacity);
  if (rc == VSRT_OK && m && events) {
    ok = cudaMalloc(&d_ev, total * sizeof(vsrt_table_event)) == cudaSuccess && (!anyhit || cudaMalloc(&d_ah, total * sizeof(vsrt_hit)) == cudaSuccess);
    tp.events = d_ev; tp.anyhit = d_ah; tp.capacity = total;
    rc = ok ? vsrt_launch_table_events(true, tp, nullptr, st) : VSRT_E_CUDA;
    if (rc == VSRT_OK && (cudaMemcpyAsync(events, d_ev, m * sizeof(vsrt_table_event), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                          (anyhit && cudaMemcpyAsync(anyhit, d_ah, m * sizeof(vsrt_hit), cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
                          cudaStreamSynchronize(st) != cudaSuccess)) rc = VSRT_E_CUDA;
  }
  cudaFree(d_tid); cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_ev); cudaFree(d_ah);
  if (rc) return fail(c, rc, "table event generation failed: %s", cudaGetErrorString(cudaGetLastError()));
  return total > capacity ? VSRT_E_CAPACITY : VSRT_OK;
}

void vsrt_table_event_stores(const vsrt_table_event* ev, uint64_t table_base, vsrt_store_txn out[2]) {
  // Baseline_warp_intersection_table::add_intersection, intersection_table.cc:180-181; Baseline_Entry = u32 hitGroupIndex[32] + {u32, u32} shader_data[32]
  const uint64_t row = table_base + (uint64_t)ev->shader_counter * 384ull;
  out[0].address = row + 4ull * ev->tid; out[0].size = 4; out[0].type = 0;            // &table[row].hitGroupIndex[tid], Intersection_Table_Store
  out[1].address = row + 128ull + 8ull * ev->tid; out[1].size = 8; out[1].type = 0;   // &table[row].shader_data[tid]
}

int vsrt_coalescing_events(vsrt_context* c, uint64_t n_rays, const uint64_t* event_offsets, const vsrt_table_event* events, vsrt_coalescing_event* out) {
  if (!c) return VSRT_E_INVALID;
  if (n_rays == 0) return VSRT_OK;
  if (!event_offsets || !out) return fail(c, VSRT_E_INVALID, "event_offsets / out is NULL");
  const uint64_t n_ev = event_offsets[n_rays];
  if (n_ev == 0) return VSRT_OK;
  if (!events) return fail(c, VSRT_E_INVALID, "events is NULL");
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  uint64_t* d_off = nullptr; vsrt_table_event* d_ev = nullptr; vsrt_coalescing_event* d_out = nullptr; uint32_t* d_err = nullptr; uint32_t h_err = 0;
  bool ok = upload(&d_off, event_offsets, n_rays + 1, st) == cudaSuccess && upload(&d_ev, events, n_ev, st) == cudaSuccess &&
            cudaMalloc(&d_out, n_ev * sizeof(vsrt_coalescing_event)) == cudaSuccess && cudaMalloc(&d_err, 4) == cudaSuccess &&
            cudaMemsetAsync(d_err, 0, 4, st) == cudaSuccess;
  int rc = ok ? vsrt_launch_coalescing(d_off, d_ev, n_rays, d_out, d_err, st) : VSRT_E_CUDA;
  if (rc == VSRT_OK && (cudaMemcpyAsync(out, d_out, n_ev * sizeof(vsrt_coalescing_event), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                        cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) rc = VSRT_E_CUDA;
  cudaFree(d_off); cudaFree(d_ev); cudaFree(d_out); cudaFree(d_err);
  if (rc) return fail(c, rc, "coalescing table replay failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (h_err & EF_UNSUPPORTED) return fail(c, VSRT_E_UNSUPPORTED, "a CTA needs more than the 100 rows the reference's Coalescing table holds");
  return VSRT_OK;
}

uint32_t vsrt_coalescing_event_stores(const vsrt_table_event* ev, const vsrt_coalescing_event* cev, uint64_t table_base, vsrt_store_txn out[3]) {
  // intersection_table.cc:73-74 (claim) / :88-90 (append); Coalescing_Entry: hitGroupIndex @0, thread_mask[32] @4, shader_data[32] @36, 292 bytes
  const uint64_t row = table_base + (uint64_t)cev->row * 292ull;
  uint32_t n = 0;
  if (cev->appended) { out[n].address = row; out[n].size = 4; out[n].type = 0; n++; }
  out[n].address = row + 4ull + ev->tid; out[n].size = 1; out[n].type = 0; n++;
  out[n].address = row + 36ull + 8ull * ev->tid; out[n].size = 8; out[n].type = 0; n++;
  return n;
}
void vsrt_coalescing_event_load(uint32_t row, uint64_t table_base, vsrt_txn* out) {
  out->address = table_base + (uint64_t)row * 292ull; out->size = 4; out->type = VSRT_TXN_INTERSECTION_TABLE_LOAD;   // :55
}

}  // extern "C"
