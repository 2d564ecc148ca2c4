/*
 * vsrt.h -- C-ABI of the B200-native functional ray-traversal path of Vulkan-Sim
 * (ubc-aamodt-group/treelet-prefetching-for-rt).
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface
 * it replaces (paths relative to the reference tree).  The library behind it is
 * hand-written sm_100a CUDA; there is no CPU fallback: without a usable CUDA device
 * vsrt_create() fails with VSRT_E_NO_DEVICE and nothing else can be called.
 *
 * Conventions: plain C types only; int return codes (0 = ok, <0 = error, never abort --
 * the reference asserts/aborts instead, vulkan_ray_tracing.cc:1568,893,2108); output
 * buffers are caller-owned with an explicit capacity and the required size is always
 * reported; a context is not thread-safe, distinct contexts are independent.
 */
#ifndef VSRT_H
#define VSRT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSRT_VERSION 1

/* ---- error codes ---- */
enum {
  VSRT_OK = 0,
  VSRT_E_INVALID = -1,       /* bad argument / bad state */
  VSRT_E_NO_DEVICE = -2,     /* no CUDA device / CUDA runtime failure at init (no CPU fallback) */
  VSRT_E_CUDA = -3,          /* CUDA error while running; see vsrt_last_error() */
  VSRT_E_CAPACITY = -4,      /* caller buffer too small; required size reported */
  VSRT_E_UNKNOWN_AS = -5,    /* TLAS/BLAS address not registered (reference: abort(), :1568) */
  VSRT_E_BAD_BVH = -6,       /* arena violates an invariant the reference asserts on */
  VSRT_E_STACK_OVERFLOW = -7,/* a ray exceeded the traversal stack capacity (raise stack_entries, at most 384) */
  VSRT_E_BUDGET = -8,        /* treelet byte budget too small (reference: assert(remaining_bytes >= 0)) */
  VSRT_E_UNSUPPORTED = -9,   /* input beyond a documented limit of this implementation: arena above 32 GiB, instance leaves of one TLAS
                                spread over more than 512 MiB, a ray with more than 4095 procedural-leaf visits, packed traces with
                                per-BLAS offsets / remap, a Coalescing table beyond the reference's 100 rows */
  VSRT_E_COMM = -10          /* NCCL not loadable or an NCCL call failed (multi-GPU counter reduce) */
};

/* ---- transaction record ABI: abstract_hardware_model.h:201-216, 315-321 ---- */
enum {
  VSRT_TXN_BVH_STRUCTURE = 0,
  VSRT_TXN_BVH_INTERNAL_NODE = 1,
  VSRT_TXN_BVH_INSTANCE_LEAF = 2,
  VSRT_TXN_BVH_PRIMITIVE_LEAF_DESCRIPTOR = 3,
  VSRT_TXN_BVH_QUAD_LEAF = 4,
  VSRT_TXN_BVH_QUAD_LEAF_HIT = 5,
  VSRT_TXN_BVH_PROCEDURAL_LEAF = 6,
  VSRT_TXN_INTERSECTION_TABLE_LOAD = 7,
  VSRT_TXN_UNDEFINED = 8
};

/* == MemoryTransactionRecord {void* address; uint32_t size; TransactionType type;} (16 B) */
typedef struct vsrt_txn {
  uint64_t address;   /* simulated-device address, host address + the offset the reference applies */
  uint32_t size;      /* 64 / 128 / 8 bytes */
  uint32_t type;      /* VSRT_TXN_* */
} vsrt_txn;

/* ---- traversal variant: instructions.cc:7235-7254 (-treelet_based_traversal) ---- */
enum {
  VSRT_MODE_DFS = 0,      /* VulkanRayTracing::traceRay, vulkan_ray_tracing.cc:2309 */
  VSRT_MODE_TREELET = 1   /* VulkanRayTracing::traceRayWithTreelets, :1522 */
};

/* ray flags honoured by the reference (SPIR-V values) */
#define VSRT_RAY_FLAG_OPAQUE 0x1u                  /* skipAnyHitShader, :2413 */
#define VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT 0x4u  /* :1650, :2411 */
#define VSRT_RAY_FLAG_SKIP_CLOSEST_HIT 0x8u        /* read, unused by traversal */

/* One trace_ray operand set: instructions.cc:7157-7227 (14 PTX operands, minus the AS handle). */
typedef struct vsrt_ray {
  float origin[3];
  float tmin;
  float direction[3];
  float tmax;
  uint32_t ray_flags;
  uint32_t cull_mask;          /* ignored by the reference traversal */
  uint32_t sbt_record_offset;  /* echoed into Traversal_data */
  uint32_t sbt_record_stride;
  uint32_t miss_index;
} vsrt_ray;                    /* 52 bytes */

/* The fields of Traversal_data / Hit_data that traversal itself writes
 * (vulkan_rt_thread_data.h:29-58; vulkan_ray_tracing.cc:2211-2245, :2990-3033). */
typedef struct vsrt_hit {
  uint32_t hit_geometry;        /* Traversal_data.hit_geometry */
  float world_min_thit;         /* closest_hit.world_min_thit */
  uint32_t primitive_index;     /* closest_hit.primitive_index  (PrimitiveIndex0) */
  uint32_t geometry_index;      /* closest_hit.geometry_index */
  uint32_t instance_index;      /* closest_hit.instance_index   (InstanceID) */
  float barycentric[3];         /* closest_hit.barycentric_coordinates {v,w,u} */
  float intersection_point[3];  /* closest_hit.intersection_point */
  uint32_t n_all_hits;          /* Traversal_data.n_all_hits (DFS mode, non-opaque rays) */
  uint64_t instance_leaf_address; /* host address of the closest instance leaf: the two 4x4 matrices of
                                     Hit_data are a pure function of its bytes (util.h:210-224) */
} vsrt_hit;                     /* 56 bytes */

/* Functional counters: cuda-sim.h:155-166, vulkan_ray_tracing.cc:1658,1687,2214,2260,2269-2282 */
typedef struct vsrt_counters {
  uint64_t mem_access_type[9];  /* g_rt_mem_access_type[TransactionType] */
  uint64_t num_hits;            /* g_rt_num_hits */
  uint64_t num_any_hits;        /* g_rt_num_any_hits */
  uint64_t n_anyhit_rays;       /* g_n_anyhit_rays (TerminateOnFirstHit rays) */
  uint64_t n_closesthit_rays;   /* g_n_closesthit_rays */
  uint64_t tot_nodes_per_ray;   /* g_tot_nodes_per_ray */
  uint64_t accessed_data_size;  /* accessedDataSize: sum of txn.size (64-bit here; the reference's is 32-bit) */
  uint64_t ray_count;           /* rayCount (global 1-based ray id of the last ray) */
  uint64_t max_nodes_per_ray;   /* g_max_nodes_per_ray */
  uint64_t max_tree_depth;      /* g_max_tree_depth */
} vsrt_counters;
#define VSRT_COUNTERS_N_SUM 16  /* leading uint64 fields reduced with SUM across ranks; the last 2 with MAX */
#define VSRT_COUNTERS_N_MAX 2

/* The RT options of gpgpusim.config this path honours (gpu-sim.cc:446-463,797,910). */
typedef struct vsrt_config {
  int32_t device;                   /* CUDA device ordinal; -1 = current device */
  uint32_t max_treelet_size;        /* -max_treelet_size (default 49152; shipped config 512) */
  uint32_t treelet_based_traversal; /* -treelet_based_traversal: default mode of vsrt_trace_ray_warp */
  uint32_t remap_to_treelet_layout; /* -remap_to_treelet_layout */
  uint32_t treelet_remap_stride;    /* -treelet_remap_stride */
  uint32_t load_treelet_metadata;   /* -load_treelet_metadata */
  uint32_t stack_entries;           /* per-ray traversal stack capacity (0 = default 96, at most 384) */
  uint32_t ray_order;               /* VSRT_RAY_ORDER_*: the order in which the GPU picks up the rays of a batch.  Results never depend
                                       on it (every output is indexed by the ray's position in the batch); it only decides which rays
                                       share a warp.  SORTED orders the batch by origin cell + direction octant (device radix sort); AUTO is the input
                                       order -- on every workload measured the sort improved cache hit rates and still slowed the traversal down */
} vsrt_config;
enum { VSRT_RAY_ORDER_AUTO = 0, VSRT_RAY_ORDER_INPUT = 1, VSRT_RAY_ORDER_SORTED = 2 };

typedef struct vsrt_context vsrt_context;

/* ---- lifetime ---- */
void vsrt_default_config(vsrt_config* cfg);
int vsrt_create(const vsrt_config* cfg, vsrt_context** out);
void vsrt_destroy(vsrt_context* ctx);
const char* vsrt_last_error(const vsrt_context* ctx);   /* ctx may be NULL: last create() failure */
/* Parse "-flag value" lines of a gpgpusim.config (option_parser.cc) into cfg; unknown flags are skipped. */
int vsrt_config_parse(vsrt_config* cfg, const char* text);

/* ---- acceleration-structure registration ----
 * Same 3-argument shape as gpgpusim_allocTLAS / gpgpusim_allocBLAS
 * (gpgpusim_calls_from_mesa.cc:168-176 -> vulkan_ray_tracing.cc:4891-4899): host address of the
 * GEN_RT_BVH header, size of the buffer that starts there, simulated-device address.  The bytes
 * are read when the next trace/form call uploads the arena (vsrt_commit), not here. */
int vsrt_alloc_tlas(vsrt_context* ctx, const void* root_addr, uint64_t buffer_size, uint64_t gpgpusim_addr);
int vsrt_alloc_blas(vsrt_context* ctx, const void* root_addr, uint64_t buffer_size, uint64_t gpgpusim_addr);
/* Upload (or re-upload after the host bytes changed) all registered buffers into the device arena and
 * validate what the reference asserts on.  Called implicitly by the first form/trace. */
int vsrt_commit(vsrt_context* ctx);

/* ---- treelet formation: VulkanRayTracing::createTreelets (:823-1470) + buildNodeToRootMap (:475) ----
 * Formed once per (context, tlas, budget); the reference forms lazily on the first ray (:1593). */
int vsrt_form_treelets(vsrt_context* ctx, const void* tlas, uint32_t max_bytes_per_treelet);
typedef struct vsrt_treelet_info {
  uint64_t n_treelets;       /* treelet_roots_addr_only.size() */
  uint64_t n_list_entries;   /* sum of de-duplicated node-list lengths */
  uint64_t n_mapped_nodes;   /* node_map_addr_only.size() */
  uint64_t total_bvh_size;   /* "Total BVH Size" the reference prints (:1364), 64-bit */
  double form_ms;            /* device time of the formation kernels */
  uint64_t scratch_bytes;    /* peak device memory the formation itself held beside its outputs (list scratch + compact list store) */
} vsrt_treelet_info;
int vsrt_treelet_info_get(vsrt_context* ctx, vsrt_treelet_info* out);
/* Treelet table in ascending root device-address order (== iteration order of the reference's std::map,
 * index == treelet_addr_to_metadata_idx).  list_offsets has n_treelets+1 entries; node_addr/node_size give
 * each treelet's node list in the reference's (BFS) order.  Any pointer may be NULL. */
int vsrt_treelet_table(vsrt_context* ctx, uint64_t* roots, uint64_t* list_offsets,
                       uint64_t* node_addr, uint32_t* node_size);
/* node_map_addr_only in ascending node device-address order. */
int vsrt_node_map(vsrt_context* ctx, uint64_t* node_addr, uint64_t* root_addr);
/* remapBVHToTreeletLayout (:1473-1509): original device address -> address in the treelet layout, ascending
 * original order; base = where the reference's gpgpusim_malloc would have put treelet_layout_bvh. */
int vsrt_treelet_remap(vsrt_context* ctx, uint64_t base, uint64_t* n_out, uint64_t* orig_addr, uint64_t* new_addr);
/* -remap_to_treelet_layout 1: every trace record carries original_bvh_to_treelet_bvh_mapping[address] and every treelet
 * id the remapped root (:1369-1412, :1682, :1763, ...).  The mapping depends on where the reference's gpgpusim_malloc
 * placed treelet_layout_bvh (:1477); pass that address here before tracing (the remap table is rebuilt when it changes). */
int vsrt_set_treelet_layout_base(vsrt_context* ctx, uint64_t base);
/* addrToTreeletID (:468): device address of the treelet root owning `addr`; VSRT_E_INVALID if unknown
 * (the reference asserts). */
int vsrt_addr_to_treelet(vsrt_context* ctx, uint64_t addr, uint64_t* root);
/* isTreeletRoot(uint8_t*) (:462): 1 / 0, <0 on error. */
int vsrt_is_treelet_root(vsrt_context* ctx, uint64_t addr);
/* treelet_addr_to_metadata_idx (:1332) */
int vsrt_treelet_metadata_idx(vsrt_context* ctx, uint64_t root, uint32_t* idx);

/* ---- traversal: traceRay / traceRayWithTreelets over a batch of lanes ----
 * Replaces the per-lane loop core_t::execute_warp_inst_t -> trace_ray_impl -> traceRay*
 * (abstract_hardware_model.cc:3052-3063, instructions.cc:7135-7254).  Ray i gets the reference's global
 * 1-based ray id rayCount+i+1 (:1665).  Outputs, all host pointers, any may be NULL:
 *   hits[n]            what traversal writes into Traversal_data
 *   trace_offsets[n+1] CSR offsets into txns / treelet_ids
 *   txns[]             thread->set_rt_transactions() contents, in the reference's order
 *   treelet_ids[]      addrToTreeletID(txn.address) (the values of the "RayID,..." line, :2256-2262, 64-bit)
 * txn_capacity counts records; *n_txn gets the number the batch produced.  If it exceeds the capacity the
 * call returns VSRT_E_CAPACITY after filling hits/trace_offsets; the trace stays on the device and
 * vsrt_trace_fetch() can still copy it out.
 * A frame-sized batch (>= 2^20 rays) is traced in windows whose results return while the next window runs; for the common
 * layout (one host span, one host->device offset, no remap) txns[] and treelet_ids[] are written by worker threads of the
 * library from 4-byte packed records, so that 4 rather than 24 bytes per record cross the PCIe link.  Environment:
 * VSRT_HOST_THREADS (workers, default 3/4 of the cores), VSRT_PIPELINE_CHUNK (rays per window, 0 = never),
 * VSRT_HOST_EXPAND=0 (copy the 16-byte records instead). */
int vsrt_trace_rays(vsrt_context* ctx, const void* tlas, int mode, uint64_t n_rays, const vsrt_ray* rays,
                    vsrt_hit* hits, uint64_t* trace_offsets, vsrt_txn* txns, uint64_t txn_capacity,
                    uint64_t* treelet_ids, uint64_t* n_txn);
int vsrt_trace_fetch(vsrt_context* ctx, vsrt_txn* txns, uint64_t txn_capacity, uint64_t* treelet_ids);
/* One warp instruction worth of lanes (n <= 32, active_mask selects lanes), mode from the config. */
int vsrt_trace_ray_warp(vsrt_context* ctx, const void* tlas, uint32_t active_mask, const vsrt_ray rays[32],
                        vsrt_hit hits[32], uint32_t txn_counts[32], vsrt_txn* txns, uint64_t txn_capacity,
                        uint64_t* n_txn);

/* Device-resident variant: rays_dev is a device pointer to n_rays vsrt_ray; results stay on the device.
 * `stream` is a cudaStream_t (NULL = the context's stream).  vsrt_trace_device_results returns the device
 * pointers of the last batch (valid until the next trace call on this context). */
int vsrt_trace_rays_device(vsrt_context* ctx, const void* tlas, int mode, uint64_t n_rays,
                           const void* rays_dev, void* stream, uint64_t* n_txn);
typedef struct vsrt_device_results {
  const void* hits;          /* vsrt_hit[n_rays] */
  const void* trace_offsets; /* uint64_t[n_rays+1] */
  const void* txns;          /* vsrt_txn[n_txn] */
  const void* treelet_ids;   /* uint32_t[n_txn]: treelet INDEX of every record (ascending-root order == treelet_addr_to_metadata_idx,
                                0xFFFFFFFF = not a BVH node); vsrt_trace_fetch expands it to the 64-bit root addresses */
  uint64_t n_rays, n_txn;
  uint64_t algorithmic_bytes; /* sum of txn.size over the batch == accessedDataSize delta */
  float traverse_ms, scan_ms, compact_ms; /* device time of the three stages of the last batch */
  uint32_t kernel_launches;  /* kernels launched by the last batch */
  float order_ms;            /* device time of the ray-order sort that preceded the traversal (0 when the batch was not sorted) */
} vsrt_device_results;
int vsrt_trace_device_results(vsrt_context* ctx, vsrt_device_results* out);

/* ---- counters and histograms ---- */
int vsrt_get_counters(vsrt_context* ctx, vsrt_counters* out);
int vsrt_reset_counters(vsrt_context* ctx);
/* Device copy of the counters as uint64[VSRT_COUNTERS_N_SUM + VSRT_COUNTERS_N_MAX] followed by the
 * per-treelet visit histogram uint64[n_treelets] (metadata-index order): the buffers a multi-GPU run
 * all-reduces (SUM over the first part and the histogram, MAX over the 2 max fields).  The pointers alias the library's
 * own buffers: reduce into copies to keep per-rank totals, or in place to make every rank hold the global ones. */
int vsrt_counters_device(vsrt_context* ctx, void** counters_dev, void** treelet_hist_dev, uint64_t* n_treelets);
/* Host copy of the per-treelet visit histogram (records whose node belongs to treelet i, metadata-index order):
 * the popularity data the RT unit's treelet prefetcher votes on (shader.cc:3424-3433), accumulated over every
 * batch since the last vsrt_reset_counters / vsrt_form_treelets. */
int vsrt_get_treelet_histogram(vsrt_context* ctx, uint64_t* hist, uint64_t capacity);

/* ---- node-visit histogram (optional) ----
 * Records per node address, as a count per 64-byte slot of the packed arena (slot -> address: vsrt_packed_layout; slot -> treelet:
 * vsrt_node_treelet_table): the per-node popularity behind the per-treelet one.  Off by default (one more kernel per batch);
 * accumulated over every batch traced while it is on, cleared by vsrt_reset_counters; included in vsrt_reduce_counters when on. */
int vsrt_enable_node_histogram(vsrt_context* ctx, int enable);
int vsrt_get_node_histogram(vsrt_context* ctx, uint64_t* visits_of_slot, uint64_t capacity, uint64_t* n_slots);

/* ---- multi-GPU: the reduce of the functional counters and the treelet visit histogram ----
 * The path shards by rays (one context per GPU, the BVH and the treelet tables replicated, rank r traces its own ray-id block);
 * nothing is exchanged on the data path.  What a multi-GPU run has to combine are the counters above (cuda-sim.h:155-166) and the
 * popularity histogram (shader.cc:3424-3433): SUM over the first VSRT_COUNTERS_N_SUM fields and the histogram, MAX over the last
 * VSRT_COUNTERS_N_MAX.  The library does it with NCCL, bound at run time (dlopen of libnccl.so.2; VSRT_NCCL_LIB overrides).
 *   vsrt_comm_unique_id   ncclGetUniqueId: rank 0 calls it and hands the 128 bytes to the other ranks by its own means
 *   vsrt_comm_init        ncclCommInitRank on the context's device (collective: every rank calls it)
 *   vsrt_comm_attach      use a communicator the caller already owns (an ncclComm_t); it is not destroyed by the library
 *   vsrt_reduce_counters  enqueue one reduce of everything this rank traced since the previous one: a snapshot kernel on `stream`
 *                         (a cudaStream_t, NULL = the context's; it must be the stream the batches ran on), then on the library's
 *                         own stream one grouped ncclAllReduce pair -- u64 header, u32 histogram deltas -- into buffers the library
 *                         owns, and a fold into the global totals.  Returns without waiting: the reduce of frame i overlaps the
 *                         traversal of frame i + 1.  Collective: every rank calls it the same number of times.
 *   vsrt_reduce_wait      make `stream` wait for the reduces enqueued so far (no host wait); stream == (void*)-1 waits on the host
 *   vsrt_reduced_get      global totals over all ranks and reduces since vsrt_comm_init / vsrt_reset_counters (host wait included).
 *                         VSRT_E_CAPACITY if 2^32 or more records were traced between two reduces (a 32-bit delta may have wrapped)
 *   vsrt_reduced_device   the device copies of the same (valid after vsrt_reduce_wait) */
#define VSRT_COMM_ID_BYTES 128
int vsrt_comm_unique_id(uint8_t id[VSRT_COMM_ID_BYTES]);
int vsrt_comm_init(vsrt_context* ctx, uint32_t n_ranks, uint32_t rank, const uint8_t id[VSRT_COMM_ID_BYTES]);
int vsrt_comm_attach(vsrt_context* ctx, void* nccl_comm, uint32_t n_ranks, uint32_t rank);
int vsrt_comm_destroy(vsrt_context* ctx);
int vsrt_reduce_counters(vsrt_context* ctx, void* stream);
int vsrt_reduce_wait(vsrt_context* ctx, void* stream);
int vsrt_reduced_get(vsrt_context* ctx, vsrt_counters* out, uint64_t* treelet_hist, uint64_t capacity);
int vsrt_reduced_device(vsrt_context* ctx, void** counters_dev, void** treelet_hist_dev, uint64_t* n_treelets);
int vsrt_reduced_get_node_histogram(vsrt_context* ctx, uint64_t* visits_of_slot, uint64_t capacity);   /* global node-visit histogram (if enabled on every rank) */

/* ---- RT-unit replay helpers: what rt_unit (gpgpu-sim/shader.cc) does with the trace, batched over the last batch ----
 * rt_unit::sort_mem_accesses (shader.cc:3012-3089) applied to every ray's list: method = -sort_method (0 strict treelet
 * order, 1 loose).  The device trace of the last batch is replaced by the sorted one (CSR offsets unchanged):
 * vsrt_trace_fetch / vsrt_trace_device_results return the sorted lists afterwards.  Sorting always starts from the
 * original order, so the call may be repeated with the other method. */
int vsrt_sort_trace(vsrt_context* ctx, int method);

/* The treelet prefetcher of rt_unit::cycle (shader.cc:3419-3640).  -treelet_prefetch_heuristic / -treelet_prefetch_threshold /
 * -load_treelet_metadata of gpgpusim.config; treelet_metadata_base = what gpgpusim_malloc returned for treelet_metadata
 * (vulkan_ray_tracing.cc:1603; the row size is (max_treelet_size / 64) * 4, :1601-1602). */
typedef struct vsrt_prefetch_config {
  uint32_t heuristic;              /* 0 always, 1 popularity threshold, 2 partial (first nodes), 3 partial (last nodes) */
  uint32_t load_treelet_metadata;
  double threshold;
  uint64_t treelet_metadata_base;
} vsrt_prefetch_config;
typedef struct vsrt_prefetch_decision {
  uint64_t treelet_root;   /* prefetched_treelet_root (device address); 0 = no thread had a pending access */
  uint32_t votes, total;   /* treelet_prefetch_priority[root], total_threads */
  uint32_t submit;         /* submit_prefetch && root != nullptr */
  uint32_t n_nodes;        /* nodes_in_treelet.size() */
  uint32_t first_node;     /* nodes [first_node, first_node + num_nodes) of the treelet's list are queued */
  uint32_t num_nodes;
} vsrt_prefetch_decision;     /* 32 bytes */
/* One vote per group of rays of the last batch (the threads of the warps resident in one RT unit).  Group g holds
 * ray_ids[group_offsets[g] .. group_offsets[g+1]) (ray_ids NULL: the ray ids themselves); ray r votes with the treelet of
 * record front[r] of its list (front NULL: the first record; a ray whose list is exhausted does not vote, :3426).  The
 * winner is the most voted treelet, lowest root address among equals (:3441).  A fresh unit per group: the caller keeps
 * last_prefetched_treelet and the queue-occupancy test (:3537,:3556). */
int vsrt_prefetch_vote(vsrt_context* ctx, const vsrt_prefetch_config* cfg, uint64_t n_groups, const uint64_t* group_offsets,
                       const uint64_t* ray_ids, const uint32_t* front, vsrt_prefetch_decision* decisions);
/* prefetch_mem_access_q entries of every decision (:3566-3620): chunk_offsets[n_groups+1] is a CSR over
 * (chunk_addr, chunk_owner) = (32-byte chunk address, address of the node or metadata row it belongs to).  Returns
 * VSRT_E_CAPACITY if capacity < *n_chunks (offsets are still complete). */
int vsrt_prefetch_chunks(vsrt_context* ctx, const vsrt_prefetch_config* cfg, uint64_t n_groups, const vsrt_prefetch_decision* decisions,
                         uint64_t* chunk_offsets, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t capacity, uint64_t* n_chunks);

/* rt_unit::schedule_next_warp (shader.cc:4307-4392): which resident warp of each RT unit issues next.  Unit u holds the
 * warps [unit_warp_offsets[u], unit_warp_offsets[u+1]) in m_current_warps order; warp w is the 32 ray ids
 * warp_ray_ids[32*w .. 32*w+31] of the last batch (~0 = no thread in that lane); stalled[w] != 0 takes a warp out
 * (NULL: none is stalled); last_prefetched[u] = the unit's last_prefetched_treelet (0 = none); front as in
 * vsrt_prefetch_vote.  scheduler = -treelet_scheduler: 0 first non-stalled warp, 1 first warp with a thread whose pending
 * access is in that treelet, 2 the warp with the most such threads; 1 and 2 fall back to 0.  pick[u] = warp index, -1 if
 * every warp is stalled. */
int vsrt_schedule_pick(vsrt_context* ctx, int scheduler, uint64_t n_units, const uint64_t* unit_warp_offsets, const uint64_t* warp_ray_ids,
                       const uint8_t* stalled, const uint64_t* last_prefetched, const uint32_t* front, int64_t* pick);

/* ---- shader-table side effects of the last batch (Baseline tables, -gpgpu_rt_intersection_table_type 0) ----
 * Every procedural-leaf visit calls intersection_table[cta]->add_intersection (vulkan_ray_tracing.cc:2171-2203 / :2951-2984)
 * and, in traceRay, every accepted triangle hit of a ray without the Opaque flag calls anyhit_table[cta]->add_intersection
 * and pushes a Hit_data (:2869-2930).  The Baseline table (intersection_table.cc:165-187) keeps one row counter per thread
 * and issues two stores per call and no loads, so the load trace is the one vsrt_trace_rays returns.  Rays
 * [32g, 32g + 32) of the batch are taken as the threads of one CTA (thread index tid_x[r], NULL = r % 32; threads that share
 * a tid_x share its row counter, in ray order); the tables are empty at the start of the batch.  The Coalescing table
 * (type 1) is derived from these events by vsrt_coalescing_events below.  The reference asserts beyond 100 rows per thread
 * (INTERSECTION_TABLE_MAX_LENGTH); no limit here. */
typedef struct vsrt_table_event {
  uint32_t table;            /* 0 intersection_table (procedural leaf), 1 anyhit_table (non-opaque triangle hit, traceRay only) */
  uint32_t shader_counter;   /* row of the table = index[tid] before the call */
  uint32_t hit_group_index;  /* InstanceContributionToHitGroupIndex of the instance leaf */
  uint32_t primitive_id;     /* PrimitiveIndex[0] of the procedural leaf / PrimitiveIndex0 of the quad leaf */
  uint32_t instance_id;      /* InstanceID */
  uint32_t tid;              /* thread index the row is written for */
  uint32_t record;           /* index of the PROCEDURAL_LEAF / QUAD_LEAF_HIT record inside the ray's trace */
  uint32_t reserved;
} vsrt_table_event;            /* 32 bytes */
typedef struct vsrt_store_txn { uint64_t address; uint32_t size; uint32_t type; } vsrt_store_txn;   /* MemoryStoreTransactionRecord, abstract_hardware_model.h:323-329 */
/* event_offsets[n_rays + 1]: CSR over events, in each ray's traversal order; anyhit (may be NULL) is parallel to events and
 * holds the Hit_data of table-1 events (world_min_thit = thit / tMult, :2895).  VSRT_E_CAPACITY if capacity < *n_events. */
int vsrt_table_events(vsrt_context* ctx, const uint8_t* tid_x, uint64_t* event_offsets, vsrt_table_event* events, vsrt_hit* anyhit,
                      uint64_t capacity, uint64_t* n_events);
/* The two MemoryStoreTransactionRecords of one event for a table that gpgpusim_alloc placed at table_base. */
void vsrt_table_event_stores(const vsrt_table_event* ev, uint64_t table_base, vsrt_store_txn out[2]);

/* ---- packed trace for host consumers ----
 * The 16-byte records + 64-bit treelet ids of a 1080p frame are 2.1 GB; a host-side caller that wants them in host memory
 * is bounded by the link or by the host's memory bandwidth (vsrt_trace_rays: ~90 M rays/s), not by the GPU.  A record is a function of 32 bits -- the node's 64-byte slot in the packed arena and a 3-bit code --
 * and a treelet id is an index into vsrt_treelet_table's ascending root array, so a caller that builds its
 * MemoryTransactionRecords where it consumes them (trace_ray_impl -> thread->set_rt_transactions, instructions.cc:7235-7254)
 * can take 8 bytes per record instead of 24 and expand with vsrt_unpack_txn.  Same records, same order, same ids.
 * Limits: every BLAS registered with the TLAS's host->device offset (the reference's two offset conventions then coincide,
 * SURVEY A.2), remap_to_treelet_layout off, at most 8 disjoint host spans; otherwise VSRT_E_UNSUPPORTED. */
typedef struct vsrt_packed_layout {
  int64_t device_delta;        /* simulated-device address - host address */
  uint32_t n_spans, reserved;
  struct { uint64_t host; uint32_t slot0, n_slots; } spans[8];   /* slot s of span i lives at spans[i].host + (s - slot0) * 64 */
} vsrt_packed_layout;
int vsrt_packed_layout_get(vsrt_context* ctx, const void* tlas, vsrt_packed_layout* out);
/* records[i] = slot << 3 | code (code = VSRT_TXN_* type, 7 = an internal node of the TLAS); treelet_index[i] = rank of the
 * record's treelet root in vsrt_treelet_table's roots[] (0xFFFFFFFF: the address is in no treelet).  Either may be NULL.
 * Always the traversal order of the last batch (vsrt_sort_trace does not affect it). */
int vsrt_trace_fetch_packed(vsrt_context* ctx, uint32_t* records, uint64_t capacity, uint32_t* treelet_index);
/* vsrt_trace_rays with packed outputs (hits and trace_offsets as there).  treelet_index == NULL is the lean form: 4 bytes per
 * record over PCIe, the treelet of a record being vsrt_node_treelet_table()[record >> 3].  In that form a frame-sized batch (at
 * least 2 x 524,288 rays; VSRT_PIPELINE_CHUNK in the environment changes the chunk, 0 disables) is traced in chunks whose
 * device->host copies overlap the upload and traversal of the next chunk; the results are the same, but there is no single
 * "last batch" afterwards for vsrt_trace_fetch* / vsrt_trace_device_results / the replay helpers to refer to. */
int vsrt_trace_rays_packed(vsrt_context* ctx, const void* tlas, int mode, uint64_t n_rays, const vsrt_ray* rays, vsrt_hit* hits,
                           uint64_t* trace_offsets, uint32_t* records, uint64_t capacity, uint32_t* treelet_index, uint64_t* n_txn);
/* Treelet index (rank of the root in vsrt_treelet_table's roots[], 0xFFFFFFFF = none) of every 64-byte slot of the packed arena:
 * addrToTreeletID (:468) for packed records, fetched once per formation.  treelet_of_slot may be NULL to query *n_slots. */
int vsrt_node_treelet_table(vsrt_context* ctx, uint32_t* treelet_of_slot, uint64_t capacity, uint64_t* n_slots);
static inline void vsrt_unpack_txn(const vsrt_packed_layout* L, uint32_t record, vsrt_txn* out) {
  const uint32_t slot = record >> 3, code = record & 7u;
  uint32_t i = L->n_spans - 1u;
  while (i > 0u && L->spans[i].slot0 > slot) i--;
  out->address = L->spans[i].host + (uint64_t)(slot - L->spans[i].slot0) * 64u + (uint64_t)L->device_delta;
  out->size = code == VSRT_TXN_BVH_INSTANCE_LEAF ? 128u : (code == VSRT_TXN_BVH_PRIMITIVE_LEAF_DESCRIPTOR ? 8u : 64u);
  out->type = code == 7u ? (uint32_t)VSRT_TXN_BVH_INTERNAL_NODE : code;
}
/* the same over an array (exported for bindings that cannot use the inline) */
void vsrt_unpack_txns(const vsrt_packed_layout* layout, const uint32_t* records, uint64_t n, vsrt_txn* out);

/* ---- Function_Call_Coalescing intersection table (-gpgpu_rt_intersection_table_type 1) ----
 * Coalescing_warp_intersection_table::add_intersection (intersection_table.cc:43-98): the rows are shared by the threads of a
 * CTA.  A call looks at rows 0, 1, ... (one Intersection_Table_Load record {&table[i].hitGroupIndex, 4} per row), claims the
 * first row of its hit group whose thread_mask[tid] is free (stores: thread_mask[tid], shader_data[tid]) or appends a row
 * (stores: hitGroupIndex, thread_mask[tid], shader_data[tid]).  traceRay / traceRayWithTreelets append the returned loads to
 * the ray's transaction list right after the PROCEDURAL_LEAF record unless the address is already in the list
 * (vulkan_ray_tracing.cc:2186-2200 / :2966-2980), i.e. only rows first_new_load .. n_loads-1.  Only the intersection table
 * has this type (the any-hit table is always a Baseline table, :441-447): events with table == 1 get an all-zero result.
 * Input: event_offsets / events as vsrt_table_events returned them (rays [32g, 32g + 32) = one CTA, in lane order, a fresh
 * table per CTA -- the reference's clear() (:101-111) cannot be followed: its inner loop increments the wrong variable and
 * never terminates on a non-empty table).  VSRT_E_UNSUPPORTED if a CTA needs more than the 100 rows the reference allocates. */
typedef struct vsrt_coalescing_event {
  uint32_t row;             /* row the thread's shader data went to */
  uint32_t appended;        /* 1: a new row was appended (three stores); 0: an existing row was claimed (two stores) */
  uint32_t n_loads;         /* Intersection_Table_Load records add_intersection returned: rows 0 .. n_loads-1 */
  uint32_t first_new_load;  /* rows first_new_load .. n_loads-1 are new to the ray's transaction list and are appended to it */
} vsrt_coalescing_event;
int vsrt_coalescing_events(vsrt_context* ctx, uint64_t n_rays, const uint64_t* event_offsets, const vsrt_table_event* events,
                           vsrt_coalescing_event* out);
/* The MemoryStoreTransactionRecords of one call (returns how many: 2 or 3) and the load record of one row, for a table that
 * gpgpusim_alloc placed at table_base (Coalescing_Entry = u32 hitGroupIndex, bool thread_mask[32], {u32, u32} shader_data[32]). */
uint32_t vsrt_coalescing_event_stores(const vsrt_table_event* ev, const vsrt_coalescing_event* cev, uint64_t table_base, vsrt_store_txn out[3]);
void vsrt_coalescing_event_load(uint32_t row, uint64_t table_base, vsrt_txn* out);

/* ---- acceleration-structure dump files (VulkanRayTracing::dump_AS, vulkan_ray_tracing.cc:4455-4558, split_files) ----
 * <prefix>.asmain / .asback / .asfront / .asmetadata: the TLAS descriptor range, the BLAS ranges below and above it and
 * the offsets findOffsetBounds (:4901) computed.  vsrt_as_dump_write produces them for a TLAS of desc_size bytes whose
 * BLAS headers are child_addrs[] (the addresses Mesa passes through gpgpusim_pass_child_addr); back/front_buffer are
 * the slack bytes written after the last backward / forward BLAS start (the reference uses 0 and 20 KiB).
 * vsrt_as_dump_read rebuilds one 64-byte aligned host image with the original relative placement (free it with
 * vsrt_as_dump_free); vsrt_register_as_image walks the TLAS in it, registers the TLAS and every BLAS an instance leaf
 * references (simulated-device address = host address + device_delta) and reports how many BLASes it found.  The TLAS
 * handle for the trace calls is image + tlas_offset. */
int vsrt_as_dump_write(const char* prefix, const void* tlas, uint64_t desc_size, const void* const* child_addrs, uint32_t n_children,
                       uint64_t back_buffer, uint64_t front_buffer);
int vsrt_as_dump_read(const char* prefix, void** image, uint64_t* image_size, uint64_t* tlas_offset);
void vsrt_as_dump_free(void* image);
int vsrt_register_as_image(vsrt_context* ctx, const void* image, uint64_t image_size, uint64_t tlas_offset, int64_t device_delta,
                           uint32_t* n_blas);

#ifdef __cplusplus
}
#endif
#endif /* VSRT_H */
