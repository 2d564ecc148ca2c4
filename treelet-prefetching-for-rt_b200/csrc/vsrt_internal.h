// vsrt_internal.h -- shared declarations of the CUDA library behind include/vsrt.h.
//
// Data layout in HBM:
//   arena        every registered AS buffer, merged into disjoint host spans, packed back to back in ascending
//                host-address order.  A node is named by its 64-byte SLOT index in the packed arena (u32).
//   node_tid     u32 per slot: index (ascending-root-address rank == the reference's treelet_addr_to_metadata_idx)
//                of the treelet that owns the node, i.e. addrToTreeletID as an index, | bit 31 when the slot is that
//                treelet's own root.  NO_TID for non-nodes.
//   root bitmap  1 bit per slot (is the slot a treelet root) + per-word exclusive popcount prefix -> rank(slot).
//   treelet CSR  tl_root[t] (slot), tl_off[t], tl_node[k] (slot | kind in the top bits of a u64).
//   staging      per ray `cap` compact trace records  (slot << 3) | code, written by the traversal kernel.
//   outputs      vsrt_hit[n], u64 offsets[n+1], vsrt_txn[total], u32 treelet index[total].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vsrt.h"

#define VSRT_NO_TID 0xFFFFFFFFu
#define VSRT_TID_SELF_ROOTED 0x80000000u   // flag in node_tid[]: the slot is the root of the treelet it is mapped to
#define VSRT_TID_MASK 0x7FFFFFFFu
#define VSRT_NO_INST 0x7FFFFFFFu
// K3 keeps a CTA-private counter for VSRT_HOT_N treelets: bucket h = index & (VSRT_HOT_N - 1) belongs to the treelet K0 found
// EARLIEST among those that map to it (formation runs in generations from the TLAS root, so earlier = nearer the top of the tree =
// visited by more rays); every other treelet of the bucket counts straight into the global histogram
#ifndef VSRT_HOT_BITS
#define VSRT_HOT_BITS 11
#endif
#define VSRT_HOT_N (1u << VSRT_HOT_BITS)
#define VSRT_MAX_SPANS_INLINE 8

// compact trace-record codes: the TransactionType (0..6) except that a TLAS internal node is tagged 7 so the
// expansion kernel can tell which host->device offset the reference would have applied in DFS mode.
enum { C_STRUCT = 0, C_INTERNAL_BLAS = 1, C_INSTANCE = 2, C_DESC = 3, C_QUAD = 4, C_QUAD_HIT = 5, C_PROC = 6, C_INTERNAL_TLAS = 7 };

// candidate / list-entry kinds of treelet formation
enum { K_TLAS_HEADER = 0, K_TLAS_INTERNAL = 1, K_INSTANCE = 2, K_BLAS_HEADER = 3, K_BLAS_INTERNAL = 4, K_BLAS_LEAF = 5 };

struct Span {             // one merged host range
  uint64_t host;          // host address of the first byte
  uint64_t size;          // bytes (multiple of 64)
  uint32_t slot0;         // first slot in the packed arena
  uint32_t n_slots;
};
struct BlasReg {          // blas_addr_map entry
  uint32_t hdr_slot;      // slot of the GEN_RT_BVH header
  uint32_t pad;
  int64_t delta;          // simulated-device address - host address
};

// Everything a kernel needs to know about the arena; passed by value.
struct ArenaView {
  const uint8_t* base;    // device pointer, 64-byte aligned
  uint32_t n_slots;
  uint32_t n_spans;
  uint32_t n_blas;
  uint32_t tlas_slot;
  const Span* spans;      // device, ascending
  const BlasReg* blas;    // device, ascending hdr_slot
  int64_t tlas_delta;     // tlas_addr - _topLevelAS
  uint32_t uniform_delta; // 1 if every BLAS delta equals tlas_delta (then both reference conventions coincide)
  uint32_t force_exact;   // 1 if some node origin is not finite: a NaN can reach the slab test, keep the ternary MIN/MAX
  uint32_t inst_base;     // lowest instance-leaf slot of the formed TLAS (K0): stack entries hold instance slots relative to it
  uint32_t pad2;
};

struct TreeletView {
  const uint32_t* node_tid;     // [n_slots]
  const uint32_t* root_bits;    // [ceil(n_slots/32)]
  const uint32_t* root_prefix;  // [ceil(n_slots/32)]
  const uint32_t* tl_root;      // [n_treelets] slot of each treelet root, ascending
  uint32_t n_treelets;
  uint32_t pad;
  const uint8_t* tnodes;        // [n_slots * 64] K1's traversal copy of the arena (internal nodes re-laid-out, the rest verbatim)
  const uint32_t* hot_keys;     // [VSRT_HOT_N]
};

struct DevCounters {            // mirrors vsrt_counters (uint64 each)
  unsigned long long v[VSRT_COUNTERS_N_SUM + VSRT_COUNTERS_N_MAX];
};
enum { CI_TYPE0 = 0, CI_NUM_HITS = 9, CI_NUM_ANY_HITS = 10, CI_N_ANYHIT_RAYS = 11, CI_N_CLOSEST_RAYS = 12, CI_TOT_NODES = 13,
       CI_ACCESSED = 14, CI_RAY_COUNT = 15, CI_MAX_NODES = 16, CI_MAX_DEPTH = 17 };

// error flags raised by kernels (OR-ed into a device word)
enum { EF_BAD_BVH = 1, EF_UNKNOWN_AS = 2, EF_STACK = 4, EF_BUDGET = 8, EF_TRACE_CAP = 16, EF_UNSUPPORTED = 32, EF_NONFINITE = 64 /* not an error */, EF_NEED_EXACT = 128 /* not an error */ };

// L1 and shared memory share 256 KB per SM, and the driver's default split gave K1 a 132 KB shared-memory carve-out for the
// 64 KB its resident CTAs use (ncu launch__shared_mem_config_size; K3 likewise, 132 for 92 KB, but K3 does not care: measured).  Asks for the smallest carve-out that holds
// `blocks_per_sm` CTAs of `func` and leaves the rest to L1.  pct_override >= 0: that percentage; -1: leave the driver's default.
template <typename F>
inline void vsrt_min_carveout(F func, int blocks_per_sm, int dev, int pct_override = -2) {
  int pct = pct_override;
  cudaFuncAttributes fa; int max_smem = 0;
  if (pct == -2 && cudaFuncGetAttributes(&fa, func) == cudaSuccess &&
      cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && max_smem > 0) {
    const size_t per_block = (fa.sharedSizeBytes + 1024 + 127) / 128 * 128;         // + the driver's 1 KB per block, 128-byte granules
    const size_t need = per_block * (size_t)blocks_per_sm;
    const size_t kb[] = { 0, 8, 16, 32, 64, 100, 132, 164, 196, 228 };              // the carve-outs sm_100 supports
    size_t pick = 228;
    for (size_t k : kb) if (k * 1024 >= need) { pick = k; break; }
    pct = (int)(pick * 1024 * 100 / (size_t)max_smem);                              // rounded down: the driver rounds a request up to the next supported size
  }
  if (pct >= 0) cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// ---- launchers (each file implements its kernels) ----
struct FormResult { uint32_t n_treelets; uint64_t n_entries; uint64_t n_mapped; uint64_t total_bvh; float ms; uint32_t nonfinite; uint32_t inst_base; uint64_t peak_scratch_bytes; };
struct FormOutputs {      // device allocations owned by the context
  uint32_t* node_tid; uint32_t* root_bits; uint32_t* root_prefix; uint32_t* tl_root; uint64_t* tl_off; uint64_t* tl_node;
  uint8_t* tnodes;        // K1's traversal copy of the arena: internal nodes in the traversal layout (treelets.cu, k_child_mask), the rest verbatim
  uint32_t* hot_keys;     // [VSRT_HOT_N] K3's CTA-private histogram table: bucket h counts treelet hot_keys[h] (NO_TID = nobody)
};
// Forms treelets for `budget`; allocates outputs with cudaMalloc (caller frees).  Returns VSRT_* code.
// `tarena` = the caller's arena-sized buffer for K1's traversal copy (FormOutputs::tnodes points at it afterwards; may be NULL).
int vsrt_launch_form_treelets(const ArenaView& av, uint32_t budget, cudaStream_t st, FormOutputs* out, FormResult* res,
                              uint32_t* err_flags_dev, char* errbuf, size_t errcap, uint8_t* tarena);

// remapBVHToTreeletLayout as a device table: remap_dev[slot] = address of the node in the treelet layout (0 = unmapped)
int vsrt_launch_remap(const FormOutputs& fo, uint32_t n_treelets, uint32_t n_slots, uint64_t base, uint64_t pitch, uint64_t* remap_dev, cudaStream_t st);

struct TraverseParams {
  ArenaView av; TreeletView tv;
  const vsrt_ray* rays; uint64_t n_rays;
  vsrt_hit* hits;
  uint32_t* stage;            // [n_rays * cap]
  uint32_t* counts;           // [n_rays] records emitted per ray
  uint32_t* nproc;            // [n_rays] procedural-leaf visits per ray (their instance refs: stage[r * cap + cap - 1 - j])
  uint32_t cap;               // staging records per ray
  uint32_t mode;
  DevCounters* counters;
  uint32_t* err_flags;
  unsigned long long* next_ray;   // global ray counter the persistent warps pull from (zeroed by the launcher)
  uint32_t refill_t;              // refill a warp when at least this many lanes are idle
  uint32_t leaf_t;                // run the leaf phase when at least this many lanes wait at a BLAS leaf
  uint32_t only_deferred;         // EXACT pass: trace only the rays the fast pass marked RAY_DEFERRED
  uint32_t gate;                  // != 0: the kernel returns at once unless (*err_flags & gate) -- lets the host queue the EXACT pass without reading the flags back
  uint2* gstack;                  // wavefront kernel: per-warp stack areas (vsrt_wf_stack_bytes)
  uint32_t stack_n;               // wavefront kernel: stack entries per ray
  uint32_t magic16;               // 0x64646464 (fp16 1024 in each half), passed as data so it lives in a register (byte_pair_f16)
  const uint32_t* perm;           // rayorder.cu: the k-th ray a lane picks up is perm[k] (NULL = input order) ...
  const uint32_t* perm_on;        // ... if this device word is non-zero (AUTO mode decides on the device)
  const uint32_t* sel;            // != NULL: the kernel returns at once unless *sel == sel_want (two hot instantiations are queued for a
  uint32_t sel_want;              //   batch, one per node layout, and k_ray_coherence's device word picks the one that runs)
};
// trav_layout: the hot kernel reads K1's traversal copy of the arena (TreeletView::tnodes) instead of the Mesa-layout arena
int vsrt_launch_traverse(const TraverseParams& p, uint32_t stack_entries, bool exact, bool trav_layout, cudaStream_t st);
// *out = 1 if the batch's consecutive rays point the same way (camera rays), 0 otherwise (bounce rays)
int vsrt_launch_ray_coherence(const vsrt_ray* rays, uint64_t n, uint32_t* out, cudaStream_t st);
// warp-wavefront formulation (traverse_wf.cu): same results, a pool of 64 rays per warp regrouped by phase every iteration
unsigned vsrt_wf_grid(uint64_t n_rays);
size_t vsrt_wf_stack_bytes(unsigned grid, uint32_t stack_entries);
int vsrt_launch_traverse_wf(const TraverseParams& p, unsigned grid, bool exact, cudaStream_t st);

// ray order for K1 (rayorder.cu): sorted ray ids + the device-side decision word, both inside `tmp` (vsrt_rayorder_tmp_bytes)
size_t vsrt_rayorder_tmp_bytes(uint64_t n);
int vsrt_launch_rayorder(const vsrt_ray* rays_dev, uint64_t n, bool force, void* tmp, const uint32_t** perm_out, const uint32_t** decision_out, cudaStream_t st);

// stable LSD radix sort of (key, id) pairs by the low 8 * passes bits (rayorder.cu); see the definition for the buffer contract
size_t vsrt_radix_tmp_bytes(uint64_t n);
int vsrt_launch_radix_sort(uint32_t* kA, uint32_t* iA, uint32_t* kB, uint32_t* iB, uint64_t n, int passes, void* tmp, const unsigned int* gate,
                           uint32_t** keys_out, uint32_t** ids_out, cudaStream_t st);

// treelet-binned wavefront K1 (traverse_tb.cu): treelet-layout copy of the arena (opaque tables), scratch size, the batch driver
int vsrt_tb_build_layout(const ArenaView& av, const FormOutputs& fo, uint32_t n_treelets, void** tables_out, cudaStream_t st);
void vsrt_tb_free_layout(void* tables);
size_t vsrt_traverse_gstack_bytes(uint32_t stack_entries);   // traverse.cu: global-memory stack of the VSRT_K1_GSTACK build, 0 otherwise
size_t vsrt_tb_scratch_bytes(uint64_t n_rays, uint32_t stack_n);
int vsrt_launch_traverse_tb(const TraverseParams& tp, void* tables, uint32_t stack_n, void* scratch, unsigned long long* stats_out, cudaStream_t st);

// exclusive scan of u32 counts into u64 offsets[n+1]; `tmp` must hold vsrt_scan_tmp_bytes(n) bytes
size_t vsrt_scan_tmp_bytes(uint64_t n);
// total_out (optional, device): also receives the grand total offsets[n]
int vsrt_launch_scan(const uint32_t* counts, uint64_t n, uint64_t* offsets, void* tmp, cudaStream_t st, unsigned long long* total_out = nullptr);

struct CompactParams {
  ArenaView av; TreeletView tv;
  const uint32_t* stage; uint32_t cap; uint32_t mode;
  const uint64_t* offsets; uint64_t n_rays;
  vsrt_txn* txns; uint32_t* tids; uint64_t out_capacity;
  uint32_t* packed;           // != NULL: write 4-byte packed records here instead of txns / tids
  uint32_t count; uint32_t pad3;   // 0: records only -- counters and histogram were already accumulated for this batch
  DevCounters* counters; unsigned long long* treelet_hist;   // may be NULL
  const uint64_t* remap;      // -remap_to_treelet_layout: record address = remap[slot] (NULL = original addresses)
  const uint32_t* err_flags;  // K3 does nothing if (*err_flags & fatal_mask) or if the batch has more records than out_capacity:
  uint32_t fatal_mask;        //   the host queues it before it has read either back, and repeats it after growing the buffers
};
int vsrt_launch_compact(const CompactParams& p, cudaStream_t st);
// optional node-visit histogram over the staged records of a batch (after the scan: offsets give the per-ray counts)
int vsrt_launch_node_hist(const uint32_t* stage, uint32_t cap, const uint64_t* offsets, uint64_t n_rays, unsigned long long* hist, const uint32_t* err_flags,
                          uint32_t fatal_mask, cudaStream_t st);

// u32 treelet index -> u64 root device address (addrToTreeletID value)
// remap_pitch != 0: roots are reported in the treelet layout, remap_base + index * remap_pitch
int vsrt_launch_tid_to_addr(const ArenaView& av, const TreeletView& tv, const uint32_t* tids, uint64_t n, uint64_t* out,
                            uint64_t remap_base, uint64_t remap_pitch, cudaStream_t st);

// ---- RT-unit replay helpers (replay.cu) ----
int vsrt_launch_build_inverse(const FormOutputs& fo, uint32_t n_treelets, uint64_t n_entries, uint32_t n_slots,
                              uint64_t** inv_off_out, uint2** inv_out, cudaStream_t st);
int vsrt_launch_sort_trace(int method, const uint64_t* offsets, uint64_t n_rays, const vsrt_txn* txns, const uint32_t* tids,
                           const uint32_t* stage, uint32_t cap, const uint64_t* inv_off, const uint2* inv,
                           vsrt_txn* out_txns, uint32_t* out_tids, uint64_t* key_scratch, cudaStream_t st);
// out_dev[g].treelet_root receives the treelet INDEX (~0 = nobody voted); the caller converts to a root address
int vsrt_launch_prefetch_vote(const uint64_t* offsets, const uint32_t* tids, uint64_t n_rays_batch, const uint64_t* group_offsets_dev,
                              const uint64_t* group_offsets_host, const uint64_t* ray_ids_dev, const uint32_t* front_dev, uint64_t n_groups,
                              const FormOutputs& fo, uint32_t n_treelets, uint32_t heuristic, double threshold,
                              vsrt_prefetch_decision* out_dev, cudaStream_t st);
int vsrt_launch_prefetch_chunks(bool fill, const ArenaView& av, const TreeletView& tv, const FormOutputs& fo, const uint64_t* remap,
                                const vsrt_prefetch_decision* dec_dev, uint64_t n_groups, uint32_t load_metadata, uint32_t per_meta, uint64_t metadata_base,
                                uint32_t* counts, const uint64_t* chunk_off, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t capacity, cudaStream_t st);
int vsrt_launch_schedule_pick(const uint64_t* offsets, const uint32_t* tids, uint64_t n_rays_batch, const uint64_t* unit_offsets_dev,
                              const uint64_t* warp_ray_ids_dev, const uint8_t* stalled_dev, const uint32_t* target_tid_dev, const uint32_t* front_dev,
                              uint64_t n_units, int scheduler, int64_t* pick_dev, cudaStream_t st);

// ---- shader-table side effects (replay.cu): Baseline warp intersection / any-hit tables, from the staged trace of the last batch
struct TableParams {
  ArenaView av;
  const vsrt_ray* rays; uint64_t n_rays;
  const uint32_t* stage; uint32_t cap; uint32_t mode;
  const uint32_t* counts; const uint32_t* nproc;
  const uint8_t* tid_x;              // [n_rays] or NULL (thread r % 32)
  uint32_t* ev_counts;               // [n_rays] events per ray: out of the count pass, in of the fill pass
  const uint64_t* ev_offsets;        // [n_rays + 1] (fill pass)
  vsrt_table_event* events; vsrt_hit* anyhit; uint64_t capacity;
};
int vsrt_launch_table_events(bool fill, const TableParams& p, uint32_t* totals_dev, cudaStream_t st);
// Function_Call_Coalescing intersection table replayed over table events (device arrays); *err receives EF_UNSUPPORTED if a CTA outgrows the reference's 100 rows
int vsrt_launch_coalescing(const uint64_t* ev_off, const vsrt_table_event* ev, uint64_t n_rays, vsrt_coalescing_event* out, uint32_t* err, cudaStream_t st);
// CSR-compacted copy of the staged trace words (vsrt_trace_fetch_packed)
int vsrt_launch_pack_trace(const uint32_t* stage, uint32_t cap, const uint64_t* offsets, uint64_t n_rays, uint32_t* out, uint64_t out_capacity, cudaStream_t st);
