// traverse.cu -- K1: per-ray functional traversal, both reference variants.
//
//   MODE_DFS      VulkanRayTracing::traceRay            (vulkan_ray_tracing.cc:2309-3076)
//   MODE_TREELET  VulkanRayTracing::traceRayWithTreelets (vulkan_ray_tracing.cc:1522-2307)
//
// One LANE owns one ray from its first record to its hit record, so the visit order is the reference's by
// construction: DFS = LIFO stack where the first box-hit internal child is followed immediately (:2573,:2760);
// TREELET = two LIFOs, children go to `current` when they live in the current treelet, else to `other`
// (:1832-1856), `other` is drained only when `current` is empty (:1748-1754).
//
// Scheduling (free to choose, because rays are independent): persistent warps.  A lane that finishes its ray is
// refilled from a global ray counter once enough lanes of the warp are idle, and each warp iteration runs three
// warp-uniform phases -- internal nodes, instance leaves, BLAS leaves -- so lanes execute the same code together.
// A lane whose next entry is a BLAS leaf WAITS (it never skips ahead: the triangle test moves min_thit, which
// culls later boxes) until leaf_t lanes are waiting or no lane has other work; this batches the triangle tests,
// which ran with ~3 of 32 lanes active in the one-thread-one-ray-one-pass formulation (profiles/README.md).
//
// EXACT = false is the hot kernel: IEEE min/max (FMNMX) in the slab test, valid only while no NaN can reach it.
// Rays (or instances) with non-finite coordinates are not traced by it: it marks them (counts[r] = RAY_DEFERRED,
// EF_NEED_EXACT) and the launcher runs the EXACT = true instantiation, which keeps the reference's ternary
// MIN/MAX, over just those rays.  An arena with a non-finite node origin runs EXACT for every ray.
//
// The kernel emits COMPACT trace records (slot << 3 | code) into the ray's staging segment; scan + K3
// (compact.cu) turn them into the reference's 16-byte MemoryTransactionRecords in CSR order.
#include "vsrt_device.cuh"
#include <algorithm>
#include <cstdlib>

#ifndef VSRT_K1_NODE_ENTRY
#define VSRT_K1_NODE_ENTRY 1   // see below
#endif
#define VSRT_K1_NODE_ENTRY_DECL VSRT_K1_NODE_ENTRY

namespace {

// stack entry, two words:
//   slot  bit 31 "root of its own treelet" (TREELET mode, consumed when the entry is taken from `other`), bit 30 leaf,
//         bits 0..28 the arena slot
//   meta  level << 23 | instance reference (23 bits).  The instance reference is the instance leaf's slot relative to the
//         first slot of the TLAS span (INST_NONE = the entry is a TLAS node).
constexpr uint32_t INST_NONE = 0x7FFFFFu;
constexpr uint32_t RAY_DEFERRED = 0xFFFFFFFFu;
#if !VSRT_K1_NODE_ENTRY_DECL
constexpr uint32_t SLOT_MASK = 0x1FFFFFFFu, SLOT_LEAF = 0x40000000u, SLOT_SELFROOT = 0x80000000u;   // per-child stack entries (VSRT_K1_NODE_ENTRY=0) only
#endif
struct Entry { uint32_t slot; uint32_t meta; };
VS_DEV uint32_t e_level(const Entry& e) { return (e.meta >> 23) & 0xffu; }
VS_DEV uint32_t e_inst(const Entry& e) { return e.meta & INST_NONE; }
VS_DEV bool e_top(const Entry& e) { return e_inst(e) == INST_NONE; }

VS_DEV uint32_t bit_index(uint32_t one_bit) { uint32_t i; asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(one_bit)); return i; }   // FLO

#ifndef VSRT_K1_THREADS
#define VSRT_K1_THREADS 128
#endif
constexpr int THREADS = VSRT_K1_THREADS;
VS_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#if defined(VSRT_K1_PF_CHILDREN) && VSRT_K1_PF_CHILDREN
VS_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#ifndef VSRT_K1_STATS
#define VSRT_K1_STATS 0   // 1: count, per inner round, how many lanes are in which state (tools/k1_lane_states.py); costs ~10 %
#endif
#if VSRT_K1_STATS
__device__ unsigned long long g_k1_depth[128];   // [d] rays whose stack (both lists) peaked at d entries, [64 + d] peak of `current` alone
__device__ unsigned long long g_k1_stats[16];   // [0] inner rounds, [1 + state] lanes in that state at the internal-node phase, [9] rounds that ran it, [10] leaf phases run, [11] lanes in them, [12] refills, [13] lanes refilled
#endif
// pop + internal-node rounds per refill / leaf vote, and the lanes that must be at an internal node for another round.
// Compile-time on purpose: as kernel parameters the same values cost 2 % (2.05 vs 2.00 ms).  Sweep (runtime knobs, ms):
// rounds 2: 2.05 (1 lane) .. 2.02 (20 lanes); 3: 2.07 .. 2.00; 4: 2.15 .. 1.99; 6: 2.35 .. 2.00
#ifndef VSRT_K1_INNER
#define VSRT_K1_INNER 2
#endif
#ifndef VSRT_K1_INT_T
#define VSRT_K1_INT_T 16   // 20 until the end of round 2; with the 192 KB L1: 16 / 20 / 24 lanes = 1.823-1.827 / 1.835-1.838 / 1.879 ms (bench), C4 bounce 2.68 / 2.70 ms
#endif
// 1 (default) = the stack holds one 16-byte entry per internal node and list -- first child slot, six offset|flag bytes, mask of the
// still pending hit children, meta -- instead of one 8-byte entry per hit child (0, kept as the A/B variant).  Both reference lists
// are LIFO and a node's hit children are pushed in slot order (:1810-1869 / :2573), so "highest pending child of the top entry" is
// the reference's pop order.  No per-child push loop (it ran 3.15 turns per node at 7.6 of 32 lanes), a slightly longer pop:
// 1.96 -> 1.88 ms on the bench workload, 3.39 -> 3.08 ms in DFS mode.  (Taking the next child straight from the registers on top
// of this was measured and is slower -- the pop then runs for a few lanes at the cost of all, profiles/README.md.)
#ifndef VSRT_K1_NODE_ENTRY
#define VSRT_K1_NODE_ENTRY 1
#endif
// L1 prefetches (prefetch.global.L1) of node bytes a lane is known to need a little later:
//   PF_LEAF  the BLAS leaf a lane has just taken -- it waits for the batched leaf phase, the leaf is cold (few rays share it)
//   PF_NEXT  the child the lane will pop after the internal node it has just tested (node-entry stack only)
// LEAF_ASYNC: the 64 bytes of a BLAS leaf are copied global -> shared memory with cp.async (no registers, no L1 line to lose)
// when the lane TAKES the leaf; the batched leaf phase, a round or two later, waits for the lane's own copies and reads them back.
// Round 2: on by default.  With the leaf threshold at 1 it costs the coherent bench workload nothing (1.923 vs 1.921 ms; traceRay mode
// 1.6 % slower) and makes the incoherent configs 7-8 % faster (C3 bounce 3.32 -> 3.08 ms, C4 3.02 -> 2.77 ms), where a leaf is a DRAM miss.
#ifndef VSRT_K1_LEAF_ASYNC
#define VSRT_K1_LEAF_ASYNC 1
#endif
// (L1 hints measured in round 2 and dropped: leaf copies that bypass L1 -- cp.async.cg -- and internal-node loads with
// L1::evict_last moved K1 by less than the run-to-run noise on all three workloads; internal-node loads that do not allocate in
// L1 cost 14 % on the headline and 10-28 % on the incoherent configs, profiles/README.md.)
#ifndef VSRT_K1_PF_LEAF
#define VSRT_K1_PF_LEAF 0
#endif
#ifndef VSRT_K1_PF_NEXT
#define VSRT_K1_PF_NEXT 0
#endif
// PF_CHILDREN: as soon as an internal node's ChildOffset has arrived, prefetch the block of its children (contiguous, at most six
// 64-byte slots) -- BEFORE the slab test decides which of them the ray visits.  Unlike PF_NEXT / PF_LEAF (issued once the child is
// known, i.e. a few dozen instructions before its demand load) this has the whole slab test, push and pop as lead time, and it
// spends DRAM bandwidth K1 does not use (5-7 % of peak on the incoherent configs, where the kernel waits for load LATENCY).
// 1 = into L2 only (prefetch.global.L2), 2 = into L1.
#ifndef VSRT_K1_PF_CHILDREN
#define VSRT_K1_PF_CHILDREN 0
#endif
#ifndef VSRT_K1_MIN_BLOCKS
#define VSRT_K1_MIN_BLOCKS 7
#endif
// 1 (default): the hot kernel traverses a second copy of the arena in which K0 has re-laid-out every internal node (treelets.cu,
// k_child_mask: absolute first-child slot, per-child offset|flag bytes, present / leaf / same-treelet masks as fields, bound bytes
// grouped so that near / far planes are whole words) instead of deriving all of that from the Mesa layout at every visit.  It is
// a full copy, leaves included, so that an internal node and its leaf siblings still share cache lines and pages: with only the
// internal nodes in a separate array the coherent headline gained 2 % and the incoherent configs LOST 12 % (two address ranges).
#ifndef VSRT_K1_TNODES
#define VSRT_K1_TNODES 1
#endif
// Shared-memory traversal stack of the hot (EXACT = false) kernel: S 16-byte node entries per lane, [entry][thread] so that a
// warp's 128-bit accesses are conflict-free whatever the lanes' depths.  0 = the stack lives in local memory (L1-cached, written
// through to L2).  A ray that needs more than S entries is handed to the EXACT pass, whose stack is the local-memory one
// (vsrt_config.stack_entries): same results, so no ray is ever refused because of S.
#ifndef VSRT_K1_SMEM_STACK
#define VSRT_K1_SMEM_STACK 0
#endif
// Traversal stack of the hot kernel in GLOBAL memory with the layout [warp][entry][lane], 16 bytes per lane.  Local memory
// interleaves the lanes of a warp word by word (word w of a lane's array lives in line w of the warp's area), so a 16-byte
// entry read at lane-dependent depths touches four sectors per lane for 16 useful bytes (ncu, round 2: 1.8 bytes per sector on the
// local loads and stores of K1, 44 % of all L1 sectors of the kernel, 115 M partial-sector writes through to L2 per launch).
// Here a lane's entry is one contiguous half-sector.  Space: resident lanes x entries x 16 B (204 MB for 96 entries).
#ifndef VSRT_K1_GSTACK
#define VSRT_K1_GSTACK 0
#endif
constexpr int PUSH_MAX = VSRT_K1_NODE_ENTRY ? 2 : 6;   // stack entries one internal node can push in TREELET mode: one per list, or one per child
enum { ST_IDLE = 0, ST_DEFER = 1, ST_FIN = 2, ST_POP = 3, ST_INT = 4, ST_INST = 5, ST_LEAF = 6 };   // lane state (DEFER: the ray is handed to the EXACT pass at the next refill)

// The ray the lane is currently testing against: the world ray inside the TLAS, the object-space ray of instance
// `inst` inside a BLAS (make_transformed_ray, :168-181).  Rebuilt only when a popped entry belongs to another
// context, so the hot loop never selects between two register sets.
struct ActiveRay { Ray8 ray; Idir idir; float tmult; uint32_t inst; bool nonfinite; };

template <int MODE, int STACK_N, bool EXACT, bool TNP>
__global__ void __launch_bounds__(THREADS, VSRT_K1_MIN_BLOCKS) k_traverse(const TraverseParams p) {
  if (p.gate && !(*reinterpret_cast<const volatile uint32_t*>(p.err_flags) & p.gate)) return;   // nothing was deferred to this pass
  if (p.sel && *reinterpret_cast<const volatile uint32_t*>(p.sel) != p.sel_want) return;        // the batch was given to the other node layout (k_ray_coherence)
  const ArenaView& av = p.av;
  const uint8_t* __restrict__ base = av.base;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // lowest instance-leaf slot of the TLAS (K0 leaves it behind the root-prefix table).  Read from memory on purpose: taken from
  // the kernel parameters (av.inst_base) the same value costs this kernel 38 bytes of spills at its 72-register budget
  const uint32_t inst_base = __ldg(p.tv.root_prefix + ((av.n_slots + 31u) >> 5));
  const int REFILL_T = (int)p.refill_t, LEAF_T = (int)p.leaf_t;
  constexpr int INT_T = VSRT_K1_INT_T, INNER_N = VSRT_K1_INNER;
  const bool only_deferred = EXACT && p.only_deferred != 0;

  // ---- functional counters (cuda-sim.h:155-166): per-CTA accumulators in shared memory, touched only when a ray is
  // finalised (keeps them out of the register file of the hot loop)
#if VSRT_K1_LEAF_ASYNC
  __shared__ uint4 s_leaf[4][THREADS];   // quarter j of the leaf of thread t (transposed: consecutive lanes, consecutive 16 bytes)
#endif
  __shared__ unsigned int s_cnt[8];   // 0 sum_nodes 1 max_nodes 2 max_level 3 n_hit 4 n_any 5 n_term 6 n_rays_done 7 err
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t max_meta = 0, err = 0;   // max_meta: the largest child meta word seen; its level field (bits 23..30) is g_max_tree_depth's candidate

  // ---- per-lane ray state.  `st` is the lane's whole control state: no ray / ray finished (hit record pending) / entry
  // wanted / an entry of one of the three kinds in `e` waiting for its phase.
  // x first child slot | y, z.lo16: per-child byte = offset (low 4 bits) | K0's flags (bits 7, 6) | z bits 16..21 pending mask | w meta
  constexpr bool SMEM = VSRT_K1_SMEM_STACK > 0 && VSRT_K1_NODE_ENTRY && !EXACT;
  constexpr bool TN = TNP && VSRT_K1_TNODES && VSRT_K1_NODE_ENTRY && !EXACT;   // internal nodes come from the traversal-layout copy
  constexpr int SN = SMEM ? VSRT_K1_SMEM_STACK : STACK_N;          // capacity of the stack this instantiation uses
#if VSRT_K1_SMEM_STACK > 0 && VSRT_K1_NODE_ENTRY
  __shared__ uint4 s_stk[EXACT ? 1 : VSRT_K1_SMEM_STACK][EXACT ? 1 : THREADS];
  uint4 l_stk[EXACT ? STACK_N : 1];
#define STK(i_) (*(SMEM ? &s_stk[(i_)][threadIdx.x] : &l_stk[(i_)]))
#elif VSRT_K1_GSTACK && VSRT_K1_NODE_ENTRY
  constexpr bool GST = !EXACT;
  uint4 l_stk[EXACT ? STACK_N : 1];
  // a 32-bit byte offset in a register, the base stays in the parameter bank: a 64-bit pointer held through the loop costs two registers and spills
  const uint32_t g_ofs = ((blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * (uint32_t)STACK_N * 32u + (unsigned)lane) * 16u;
#define STK(i_) (*(GST ? reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.gstack) + (g_ofs + (uint32_t)(i_) * 512u)) : &l_stk[(i_)]))
#elif VSRT_K1_NODE_ENTRY
  uint4 stk[STACK_N];
#define STK(i_) stk[(i_)]
#else
  Entry stk[STACK_N];
#define STK(i_) stk[(i_)]
#endif
  uint32_t st = ST_IDLE; bool exhausted = false;
  Entry e; e.slot = 0; e.meta = 0;
  uint32_t r = 0;
  uint32_t* rstage = p.stage;            // staging segment of ray r
  float w_tmin = 0.0f, w_tmax = 0.0f;   // the world ray's origin/direction are re-read from p.rays[r] when needed
  ActiveRay a; a.ray.ox = a.ray.oy = a.ray.oz = a.ray.dx = a.ray.dy = a.ray.dz = a.ray.tmin = a.ray.tmax = 0.0f; a.idir.x = a.idir.y = a.idir.z = 0.0f; a.tmult = 1.0f; a.inst = INST_NONE; a.nonfinite = false;
  // ray_nodes: bits 0..19 total_nodes_accessed, bits 20..31 procedural-leaf visits (their instance refs sit at the END of
  // the ray's staging segment, visit j at [cap - 1 - j], for the shader-table post-pass)
  uint32_t flags = 0, cnt = 0, ray_nodes = 0, ray_any = 0;
#if VSRT_K1_STATS
  int peak_all = 0, peak_cur = 0;
#define STAT_DEPTH() do { peak_all = max(peak_all, cur_n + oth_n); peak_cur = max(peak_cur, cur_n); } while (0)
#else
#define STAT_DEPTH() do { } while (0)
#endif
  // TREELET: the current treelet (current_treelet_root, :1707/:1752) is kept lazily.  tid_known: cur_tid is its index;
  // otherwise cur_tid holds the SLOT of the self-rooted node that was moved over from `other`, and the index is looked up
  // only when something has to be compared with it (BLAS roots, the rare generic path).
  uint32_t cur_tid = VSRT_NO_TID; bool tid_known = true;
  int cur_n = 0, oth_n = 0;
  bool in_cur = false;   // TREELET: the node in `e` is known to belong to the current treelet (node_tid == current index)
  float min_thit = 0.0f, min_thit_object = 0.0f;
  uint32_t closest_leaf = 0, closest_inst = INST_NONE;
  const uint32_t cap = p.cap;

// A/B: staged records are written once and read once, by K3 -- streaming stores (evict-first) were meant to keep them from
// displacing node bytes in L2, and are slower
#ifndef VSRT_K1_EMIT_STREAMING
#define VSRT_K1_EMIT_STREAMING 0   // measured: 1.887 vs 1.855 ms with st.global.cs
#endif
#if VSRT_K1_EMIT_STREAMING
#define EMIT(slot_, code_) do { if (cnt < cap) __stcs(rstage + cnt, ((slot_) << 3) | (uint32_t)(code_)); cnt++; } while (0)
#else
#define EMIT(slot_, code_) do { if (cnt < cap) rstage[cnt] = ((slot_) << 3) | (uint32_t)(code_); cnt++; } while (0)
#endif
#define PUSH_CUR(c_) do { STK(cur_n) = (c_); cur_n++; } while (0)
#define PUSH_OTH(c_) do { oth_n++; STK(SN - oth_n) = (c_); } while (0)
  // no room for what one node can push: an error for the local-memory stack, a hand-over to the EXACT pass for the shared one
#define STACK_FULL() do { if (SMEM) st = ST_DEFER; else err |= EF_STACK; } while (0)
#define CUR_TID() (tid_known ? cur_tid : (tid_known = true, cur_tid = __ldg(p.tv.node_tid + cur_tid) & VSRT_TID_MASK))
  // switch the active ray to context `inst_` (INST_NONE = world)
#define LOAD_WORLD(w_) do { const vsrt_ray* rp_ = p.rays + r; \
      (w_).ox = __ldg(&rp_->origin[0]); (w_).oy = __ldg(&rp_->origin[1]); (w_).oz = __ldg(&rp_->origin[2]); (w_).tmin = w_tmin; \
      (w_).dx = __ldg(&rp_->direction[0]); (w_).dy = __ldg(&rp_->direction[1]); (w_).dz = __ldg(&rp_->direction[2]); (w_).tmax = w_tmax; } while (0)
#define ACTIVATE(inst_) do { const uint32_t i_ = (inst_); if (a.inst != i_) { a.inst = i_; Ray8 w_; LOAD_WORLD(w_); \
      if (i_ == INST_NONE) { a.ray = w_; a.idir = calc_idir(w_); a.tmult = 1.0f; a.nonfinite = false; } \
      else { InstCtx c_; make_object_ray(base, inst_base + i_, w_, c_); a.ray = c_.ray; a.idir = c_.idir; a.tmult = c_.tmult; a.nonfinite = c_.exact; } } } while (0)

  while (true) {
    // ================= refill: finalize finished rays, fetch new ones
    const unsigned idle = __ballot_sync(full, st <= ST_FIN);
    if (idle == full || (!exhausted && __popc(idle) >= REFILL_T)) {
      if (st == ST_DEFER) { p.counts[r] = RAY_DEFERRED; err |= EF_NEED_EXACT; st = ST_IDLE; }   // degenerate instance transform: left to the EXACT pass
      if (st == ST_FIN) {
        // ---- hit record (:2211-2245 / :2990-3033) and per-ray counters
        st = ST_IDLE;
        if (cnt + (ray_nodes >> 20) > cap) err |= EF_TRACE_CAP;       // records and procedural-visit words share the segment
        p.counts[r] = cnt;
        p.nproc[r] = ray_nodes >> 20; ray_nodes &= 0xFFFFFu;
        vsrt_hit h;
        h.hit_geometry = 0; h.world_min_thit = 0.0f; h.primitive_index = 0; h.geometry_index = 0; h.instance_index = 0;
        h.barycentric[0] = h.barycentric[1] = h.barycentric[2] = 0.0f;
        h.intersection_point[0] = h.intersection_point[1] = h.intersection_point[2] = 0.0f;
        h.n_all_hits = ray_any; h.instance_leaf_address = 0;
        if (min_thit < w_tmax) {
          atomicAdd(&s_cnt[3], 1u);
          Ray8 w; LOAD_WORLD(w);
          const Node64 q = load_node(base, closest_leaf);
          ACTIVATE(closest_inst);
          const uint32_t ci = inst_base + closest_inst;
          h.hit_geometry = 1; h.world_min_thit = min_thit;
          h.geometry_index = q.w[1] & 0x0fffffffu; h.primitive_index = q.w[2];
          h.instance_index = __ldg(reinterpret_cast<const uint32_t*>(base + (uint64_t)ci * 64u + 72u));
          h.intersection_point[0] = fadd(w.ox, fmul(w.dx, min_thit));
          h.intersection_point[1] = fadd(w.oy, fmul(w.dy, min_thit));
          h.intersection_point[2] = fadd(w.oz, fmul(w.dz, min_thit));
          barycentric(q, fadd(a.ray.ox, fmul(a.ray.dx, min_thit_object)), fadd(a.ray.oy, fmul(a.ray.dy, min_thit_object)),
                      fadd(a.ray.oz, fmul(a.ray.dz, min_thit_object)), h.barycentric);
          h.instance_leaf_address = slot_to_host(av, ci);
        }
        p.hits[r] = h;
        atomicAdd(&s_cnt[0], ray_nodes); atomicMax(&s_cnt[1], ray_nodes);
        if (ray_any) atomicAdd(&s_cnt[4], ray_any);
        if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) atomicAdd(&s_cnt[5], 1u);
        atomicAdd(&s_cnt[6], 1u);
#if VSRT_K1_STATS
        atomicAdd(&g_k1_depth[min(peak_all, 63)], 1ull); atomicAdd(&g_k1_depth[64 + min(peak_cur, 63)], 1ull); peak_all = peak_cur = 0;
#endif
      }
#if VSRT_K1_STATS
      if (lane == 0) { atomicAdd(&g_k1_stats[12], 1ull); atomicAdd(&g_k1_stats[13], (unsigned long long)__popc(idle)); }
#endif
      if (!exhausted) {
        const int n_idle = __popc(idle);
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(p.next_ray, (unsigned long long)n_idle);
        b0 = __shfl_sync(full, b0, 0);
        if (b0 + (unsigned long long)n_idle >= p.n_rays) exhausted = true;
        if (st == ST_IDLE) {
          const uint64_t k = b0 + (uint64_t)__popc(idle & ((1u << lane) - 1u));
          // rays are handed out in sorted order when rayorder.cu decided so (the decision word is re-read here rather than kept in
          // a register of the hot loop); every output stays indexed by the ray's own id
          const uint64_t nr = (k < p.n_rays && p.perm != nullptr && __ldg(p.perm_on) != 0u) ? (uint64_t)__ldg(p.perm + k) : k;
          if (k < p.n_rays && (!only_deferred || p.counts[nr] == RAY_DEFERRED)) {
            // ---- start ray nr (:1650-1741 / :2411-2484)
            r = (uint32_t)nr; rstage = p.stage + nr * cap;
            const vsrt_ray* rp = p.rays + r;
            Ray8 w;
            w.ox = __ldg(&rp->origin[0]); w.oy = __ldg(&rp->origin[1]); w.oz = __ldg(&rp->origin[2]); w.tmin = __ldg(&rp->tmin);
            w.dx = __ldg(&rp->direction[0]); w.dy = __ldg(&rp->direction[1]); w.dz = __ldg(&rp->direction[2]); w.tmax = __ldg(&rp->tmax);
            flags = __ldg(&rp->ray_flags);
            if (!EXACT && ray_needs_exact(w)) { p.counts[r] = RAY_DEFERRED; err |= EF_NEED_EXACT; }   // left to the EXACT pass
            else {
              st = ST_POP;
              w_tmin = w.tmin; w_tmax = w.tmax;
              a.inst = INST_NONE; a.ray = w; a.idir = calc_idir(w); a.tmult = 1.0f; a.nonfinite = false;
              cnt = 0; ray_nodes = 0; ray_any = 0;
              cur_n = 0; oth_n = 0; cur_tid = VSRT_NO_TID; tid_known = true;
              min_thit = w_tmax; min_thit_object = 0.0f; closest_leaf = 0; closest_inst = INST_NONE;   // :1671
              EMIT(av.tlas_slot, C_STRUCT);                                                // :1685 / :2447
              uint32_t top_root = 0;
              if (!header_root(av, av.tlas_slot, top_root)) err |= EF_BAD_BVH;
              else {
                // scene box of the TLAS header (:1712-1741 / :2472-2484)
                const float* hb = reinterpret_cast<const float*>(base + (uint64_t)av.tlas_slot * 64u + 8u);
                float th;
                if (ray_box(__ldg(hb), __ldg(hb + 1), __ldg(hb + 2), __ldg(hb + 3), __ldg(hb + 4), __ldg(hb + 5), a.idir, w, th)) {
#if VSRT_K1_NODE_ENTRY
                  uint4 c = make_uint4(top_root, 0u, 1u << 16, (1u << 23) | INST_NONE);   // a one-child entry
#define SET_SELFROOT(c_) ((c_).y = 0x80u)
#else
                  Entry c; c.slot = top_root; c.meta = (1u << 23) | INST_NONE;
#define SET_SELFROOT(c_) ((c_).slot |= SLOT_SELFROOT)
#endif
                  if (MODE == VSRT_MODE_TREELET) {
                    cur_tid = root_rank(p.tv, av.tlas_slot);                               // current_treelet_root = TLAS + offset, :1707
                    const uint32_t tr = __ldg(p.tv.node_tid + top_root);
                    if ((tr & VSRT_TID_MASK) == cur_tid) PUSH_CUR(c); else { if (tr & VSRT_TID_SELF_ROOTED) SET_SELFROOT(c); PUSH_OTH(c); }
                  } else PUSH_CUR(c);
                  if (max_meta < (1u << 23)) max_meta = 1u << 23;
                }
              }
            }
          }
        }
      }
      if (__ballot_sync(full, st > ST_FIN) == 0u) { if (exhausted) break; else continue; }
    }

#if VSRT_K1_LEAF_ASYNC == 2
// predicated form: the four copies are issued by every popping lane under a predicate instead of inside a branch that a lane or
// two of the warp takes -- the same dozen instructions per pop round whether one lane took a leaf or twenty
#define LEAF_FETCH() do { \
      const uint8_t* g_ = base + (uint64_t)e.slot * 64u; \
      const uint32_t sa_ = (uint32_t)__cvta_generic_to_shared(&s_leaf[0][threadIdx.x]); \
      asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %0, %1;\n" \
                   " @p cp.async.ca.shared.global [%2], [%3], 16;\n @p cp.async.ca.shared.global [%2+%4], [%3+16], 16;\n" \
                   " @p cp.async.ca.shared.global [%2+%5], [%3+32], 16;\n @p cp.async.ca.shared.global [%2+%6], [%3+48], 16;\n}" \
                   ::"r"(st), "r"((uint32_t)ST_LEAF), "r"(sa_), "l"(g_), "n"(THREADS * 16), "n"(THREADS * 32), "n"(THREADS * 48) : "memory"); \
      asm volatile("cp.async.commit_group;" ::: "memory"); } while (0)
#elif VSRT_K1_LEAF_ASYNC
#define LEAF_FETCH() do { if (st == ST_LEAF) { \
      const uint8_t* g_ = base + (uint64_t)e.slot * 64u; \
      _Pragma("unroll") for (int j_ = 0; j_ < 4; j_++) { \
        const uint32_t sa_ = (uint32_t)__cvta_generic_to_shared(&s_leaf[j_][threadIdx.x]); \
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa_), "l"(g_ + 16 * j_) : "memory"); } \
      asm volatile("cp.async.commit_group;" ::: "memory"); } } while (0)
#else
#define LEAF_FETCH() do { } while (0)
#endif
    // ================= pop the next entry of every lane that wants one
    // TAKE: `e` holds the entry just taken from `current` (from_cur_) or from the front of `other`
#define TAKE(from_cur_, selfroot_in_, leaf_in_) do { \
      const bool selfroot_ = (selfroot_in_), leaf_ = (leaf_in_); \
      if (MODE == VSRT_MODE_TREELET) { \
        if (from_cur_) { if (!VSRT_K1_NODE_ENTRY) cur_n--; in_cur = true; }       /* entries of `current` were pushed because node_tid == current treelet */ \
        else { \
          /* :1748-1754 -- the front of `other` moves to `current`; current_treelet_root becomes that node's HOST address, */ \
          /* which equals a treelet's device address only for the root at (host - tlas_delta). */ \
          if (!VSRT_K1_NODE_ENTRY) oth_n--; \
          if (av.tlas_delta == 0) { \
            in_cur = selfroot_; cur_tid = e.slot; tid_known = false; \
            if (!selfroot_) { cur_tid = root_rank(p.tv, e.slot); tid_known = true; } \
          } else { uint32_t s2_; in_cur = false; tid_known = true; cur_tid = host_to_slot(av, slot_to_host(av, e.slot) - (uint64_t)av.tlas_delta, s2_) ? root_rank(p.tv, s2_) : VSRT_NO_TID; } \
        } \
      } else if (!VSRT_K1_NODE_ENTRY) cur_n--; \
      st = !leaf_ ? ST_INT : (e_top(e) ? ST_INST : ST_LEAF); \
      if (VSRT_K1_PF_LEAF && st == ST_LEAF) prefetch_l1(base + (uint64_t)e.slot * 64u); \
      LEAF_FETCH(); } while (0)
    // pop + internal-node phase run up to INNER_N times back to back (VSRT_K1_INNER): the refill and leaf votes around them are
    // amortised, at the price of idle / leaf lanes waiting a little longer
#pragma unroll 1
    for (int inner = 0; inner < INNER_N; inner++) {
    if (st == ST_POP) {
      const bool fc = cur_n != 0;
      if (fc || (MODE == VSRT_MODE_TREELET && oth_n != 0)) {
#if VSRT_K1_NODE_ENTRY
        const int idx = fc ? cur_n - 1 : SN - oth_n;
        const uint4 t = STK(idx);
        uint32_t ci; asm("bfind.u32 %0, %1;" : "=r"(ci) : "r"(t.z));          // highest pending child: the mask is the top of z
        const uint32_t z2 = t.z ^ (1u << ci);
#ifndef VSRT_K1_POP_STORE_ALWAYS
#define VSRT_K1_POP_STORE_ALWAYS 0   // 1: write the mask back unconditionally (no predicate): 1.864 vs 1.855 ms
#endif
        if (VSRT_K1_POP_STORE_ALWAYS || (z2 >> 16)) STK(idx).z = z2;               // a removed entry is never read again
        const int gone = (z2 >> 16) == 0u ? 1 : 0;
        cur_n -= fc ? gone : 0;
        if (MODE == VSRT_MODE_TREELET) oth_n -= fc ? 0 : gone;
        const uint32_t cb = __byte_perm(t.y, t.z, 0x7770u + (ci - 16u));          // the child's byte: offset | flags
        e.slot = t.x + (cb & 15u); e.meta = t.w;
        TAKE(fc, (cb & 0x80u) != 0u, (cb & 0x40u) != 0u);
#else
        e = STK(fc ? cur_n - 1 : SN - oth_n);
        const uint32_t fl = e.slot; e.slot &= SLOT_MASK;
        TAKE(fc, (fl & SLOT_SELFROOT) != 0u, (fl & SLOT_LEAF) != 0u);
#endif
      } else st = ST_FIN;
    }
    if (inner && __popc(__ballot_sync(full, st == ST_INT)) < INT_T) break;   // too few lanes at an internal node: let the other phases / the refill bring lanes back first

#if VSRT_K1_STATS
    {
      unsigned long long add[16] = { 0 }; add[0] = 1;
      for (uint32_t sv = 0; sv <= ST_LEAF; sv++) add[1 + sv] = __popc(__ballot_sync(full, st == sv));
      add[9] = __any_sync(full, st == ST_INT) ? 1 : 0;
      if (lane < 10 && add[lane]) atomicAdd(&g_k1_stats[lane], add[lane]);
    }
#endif
    // ================= phase 1: internal nodes (TLAS :1759-1875 / :2500-2599, BLAS :1954-2072 / :2687-2786)
    if (st == ST_INT) {
      st = ST_POP;
      // the hot kernel runs over the traversal copy of the arena (the launcher put it into av.base): internal nodes are in the
      // layout K0 prepared (TN), everything else is the arena's own bytes; the EXACT pass and the other K1 variants read the arena
      const Node64 n = load_node(base, e.slot);
#if VSRT_K1_PF_CHILDREN
      {
        const uint8_t* cb_ = base + (uint64_t)(TN ? n.w[3] : e.slot + (uint32_t)node_child_offset(n)) * 64u;
        if (VSRT_K1_PF_CHILDREN == 2) { prefetch_l1(cb_); prefetch_l1(cb_ + 128); prefetch_l1(cb_ + 256); }
        else { prefetch_l2(cb_); prefetch_l2(cb_ + 128); prefetch_l2(cb_ + 256); }
      }
#endif
      const uint32_t inst = e_inst(e);
      EMIT(e.slot, inst == INST_NONE ? C_INTERNAL_TLAS : C_INTERNAL_BLAS); ray_nodes++;
      ACTIVATE(inst);
      if (!EXACT && a.nonfinite) st = ST_DEFER;       // degenerate instance transform
      else {
        const float cull = fmul(min_thit, a.tmult);                                     // :1791 / :1989 (tMult is 1 in the TLAS)
        uint32_t mask, child0, lo4 = 0, hi2 = 0, xlo = 0, xhi = 0;
        uint32_t ey, ez;       // byte i of {ey, ez.lo16} = offset of child i from the first child | K0's two flags (bit 7 self-rooted, bit 6 leaf)
        if (TN) { mask = test_children_t(n, a.ray, a.idir, cull, p.magic16); child0 = n.w[3]; ey = n.w[4]; ez = n.w[5] & 0xffffu; }
        else {
          mask = test_children<EXACT>(n, a.ray, a.idir, cull, p.magic16);
          // child i lives at first_child + sum_{j<i} ChildSize[j] (:1868).  The six info bytes (22..27) are handled as packed
          // bytes: prefix sums of the sizes by one multiply; K0 left two flags in the bits the reference ignores (& 0x3f):
          // bit 7 "this child is the root of its own treelet", bit 6 "leaf" (type != 0) -- exactly bits 31/30 of an entry's slot
          lo4 = __byte_perm(n.w[5], n.w[6], 0x5432); hi2 = n.w[6] >> 16;
          const uint32_t pre4 = (lo4 & 0x03030303u) * 0x01010101u;                       // byte j = size_0 + .. + size_j
          const uint32_t t4 = pre4 >> 24; xlo = pre4 << 8; xhi = t4 | ((t4 + (hi2 & 3u)) << 8);   // byte i of {xlo, xhi} = offset of child i
          child0 = e.slot + (uint32_t)node_child_offset(n);
          ey = xlo | (lo4 & 0xC0C0C0C0u); ez = xhi | (hi2 & 0xC0C0u);
        }
        // children are one level down; the 8-bit level saturates at 255: a carry out of the field lands in bit 31 and is taken back
        const uint32_t cm_ = e.meta + (1u << 23), cmeta = cm_ - ((cm_ >> 31) << 23);
        if (mask && cmeta > max_meta) max_meta = cmeta;
        // hit children in slot order (:1810-1869): one loop turn per hit child (a node rarely has more than two)
        if (MODE == VSRT_MODE_TREELET) {
          // which children belong to the CURRENT treelet (:1832): K0 left "child i is in this node's treelet" in the
          // node's pad byte (+17), valid whenever the node itself is in the current treelet; otherwise (a node taken from
          // `other` that is not the root of its treelet, or the host/device offset quirk) look the children up
          uint32_t mc = TN ? n.w[6] >> 24 : node_byte(n, 17);
          if (!in_cur) {
            mc = 0;
            const uint32_t ct = CUR_TID();
            for (uint32_t m = mask; m; m &= m - 1u) {
              const uint32_t sel = 0x7770u + bit_index(m & (0u - m));
              if ((__ldg(p.tv.node_tid + child0 + (__byte_perm(ey, ez, sel) & 15u)) & VSRT_TID_MASK) == ct) mc |= m & (0u - m);
            }
          }
          const uint32_t mcur = mask & mc;
          if (cur_n + oth_n + PUSH_MAX > SN) STACK_FULL();      // room for everything this node can push
#if VSRT_K1_NODE_ENTRY
          else {
            const uint32_t moth = mask ^ mcur;
            if (moth) { oth_n++; STK(SN - oth_n) = make_uint4(child0, ey, ez | (moth << 16), cmeta); }
            if (mcur) { STK(cur_n) = make_uint4(child0, ey, ez | (mcur << 16), cmeta); cur_n++; }
            STAT_DEPTH();
#if VSRT_K1_PF_NEXT
            {   // the child this lane pops next, if it is one of this node's: highest hit child in `current`, else (nothing older in `current`) in `other`
              const uint32_t mnext = mcur ? mcur : (cur_n == 0 ? moth : 0u);
              if (mnext) { uint32_t ti; asm("bfind.u32 %0, %1;" : "=r"(ti) : "r"(mnext)); prefetch_l1(base + (uint64_t)(child0 + (__byte_perm(ey, ez, 0x7770u + ti) & 15u)) * 64u); }
            }
#endif
          }
#else
          else {
            int po = SN - 1 - oth_n;
            for (uint32_t m = mask; m; ) {
              const uint32_t bit = m & (0u - m); m ^= bit;
              const uint32_t sel = 0x7770u + bit_index(bit);
              Entry c; c.slot = (child0 + __byte_perm(xlo, xhi, sel)) | ((__byte_perm(lo4, hi2, sel) << 24) & 0xC0000000u); c.meta = cmeta;
              const bool ic = (mcur & bit) != 0u;
              STK(ic ? cur_n : po) = c;
              if (ic) cur_n++; else po--;
            }
            oth_n = SN - 1 - po;
            // (taking the entry just pushed from the registers instead of the stack load of the next pop -- the top stall of
            // the kernel -- was measured twice and is slower: more instructions in divergent code, profiles/README.md)
          }
#endif
        } else {
          // the first hit internal child is followed at once (:2573); every other hit child is pushed in slot order
          if (cur_n + (VSRT_K1_NODE_ENTRY ? 1 : 6) > SN) STACK_FULL();
#if VSRT_K1_NODE_ENTRY
          else if (mask) {
            uint32_t leaf6;                                                                   // leaf flag of every child -> six bits
            if (TN) leaf6 = n.w[5] >> 24;
            else {
              const uint32_t lf4 = (lo4 >> 6) & 0x01010101u, lf2 = (hi2 >> 6) & 0x0101u;
              leaf6 = (((lf4 * 0x00204081u) >> 21) & 15u) | (((lf2 * 0x00204081u) >> 17) & 0x30u);
            }
            const uint32_t mi = mask & ~leaf6, first = mi & (0u - mi), rest = mask ^ first;
            if (first) { e.slot = child0 + (__byte_perm(ey, ez, 0x7770u + bit_index(first)) & 15u); e.meta = cmeta; st = ST_INT; }
            if (rest) { STK(cur_n) = make_uint4(child0, ey, ez | (rest << 16), cmeta); cur_n++; }
            STAT_DEPTH();
          }
#else
          else {
            for (uint32_t m = mask; m; ) {
              const uint32_t bit = m & (0u - m); m ^= bit;
              const uint32_t sel = 0x7770u + bit_index(bit);
              const uint32_t fl = (__byte_perm(lo4, hi2, sel) << 24) & 0xC0000000u;
              Entry c; c.slot = child0 + __byte_perm(xlo, xhi, sel); c.meta = cmeta;
              if (!(fl & SLOT_LEAF) && st != ST_INT) { e = c; st = ST_INT; }
              else { c.slot |= fl; STK(cur_n) = c; cur_n++; }
            }
          }
#endif
        }
      }
    }
    }   // inner
    // ================= phase 2: instance leaves (:1876-1953 / :2602-2677)
    if (st == ST_INST) {
      st = ST_POP;
      EMIT(e.slot, C_INSTANCE); ray_nodes++;
      uint32_t hdr = 0, broot = 0;
      const uint32_t iref = e.slot - inst_base;
      if (!instance_blas_header(av, e.slot, hdr) || !header_root(av, hdr, broot)) { err |= EF_BAD_BVH; st = ST_FIN; }
      else if (e.slot < inst_base || iref >= INST_NONE) { err |= EF_BAD_BVH; st = ST_FIN; }   // cannot happen for an arena K0 accepted (it bounds the instance-leaf span)
      else {
        EMIT(hdr, C_STRUCT);                                                     // BLAS header record, :1913 / :2645
#if VSRT_K1_NODE_ENTRY
        uint4 c = make_uint4(broot, 0u, 1u << 16, (e_level(e) << 23) | iref);
#else
        Entry c; c.slot = broot; c.meta = (e_level(e) << 23) | iref;             // BLAS root inherits the leaf's level (:1944)
#endif
        if (MODE == VSRT_MODE_DFS) { if (cur_n < SN) PUSH_CUR(c); else STACK_FULL(); }
        else {
          const uint32_t tb = __ldg(p.tv.node_tid + broot);
          if (cur_n + oth_n >= SN) STACK_FULL();
          else if ((tb & VSRT_TID_MASK) == CUR_TID()) PUSH_CUR(c);
          else { if (tb & VSRT_TID_SELF_ROOTED) SET_SELFROOT(c); PUSH_OTH(c); }
        }
      }
    }

    // ================= phase 3: BLAS leaves (:2073-2204 / :2789-2985), batched
    const unsigned m_leaf = __ballot_sync(full, st == ST_LEAF);
    if (m_leaf && (__popc(m_leaf) >= LEAF_T || __ballot_sync(full, st == ST_INT || st == ST_INST || st == ST_POP) == 0u)) {
#if VSRT_K1_STATS
      if (lane == 0) { atomicAdd(&g_k1_stats[10], 1ull); atomicAdd(&g_k1_stats[11], (unsigned long long)__popc(m_leaf)); }
#endif
      if (st == ST_LEAF) {
        st = ST_POP;
#if VSRT_K1_LEAF_ASYNC
        Node64 q;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        { const uint4 a0 = s_leaf[0][threadIdx.x], a1 = s_leaf[1][threadIdx.x], a2 = s_leaf[2][threadIdx.x], a3 = s_leaf[3][threadIdx.x];
          q.w[0] = a0.x; q.w[1] = a0.y; q.w[2] = a0.z; q.w[3] = a0.w; q.w[4] = a1.x; q.w[5] = a1.y; q.w[6] = a1.z; q.w[7] = a1.w;
          q.w[8] = a2.x; q.w[9] = a2.y; q.w[10] = a2.z; q.w[11] = a2.w; q.w[12] = a3.x; q.w[13] = a3.y; q.w[14] = a3.z; q.w[15] = a3.w; }
#else
        const Node64 q = load_node_now(base, e.slot);
#endif
        EMIT(e.slot, C_DESC);
        if (((q.w[1] >> 29) & 1u) == 0u) {
          ACTIVATE(e_inst(e));
          float thit = 0.0f;
          const bool hit = ray_tri(q, a.ray, thit);
          // world t = t_obj / tMult (:2124 / :2840), needed only for a hit; x / 1.0f is x, and a miss leaves thit = 0, which
          // would send the IEEE division down its slow path for nothing
          const float tw = !hit ? 0.0f : (a.tmult == 1.0f ? thit : fdiv(thit, a.tmult));
          bool acc = hit && w_tmin <= tw && tw <= w_tmax;                         // :2843
          if (MODE == VSRT_MODE_TREELET) acc = acc && tw < min_thit;              // :2127
          if (acc) {
            if (MODE == VSRT_MODE_TREELET) min_thit = tw;
            else {
              const bool opaque = (flags & VSRT_RAY_FLAG_OPAQUE) != 0;            // skipAnyHitShader, :2413
              if (opaque && tw < min_thit) min_thit = tw;                         // :2850
              if (!opaque) ray_any++;                                             // :2869-2929
            }
            min_thit_object = thit; closest_leaf = e.slot; closest_inst = e_inst(e);
            EMIT(e.slot, C_QUAD_HIT); ray_nodes++;
            if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) { cur_n = 0; oth_n = 0; }   // :2151-2155 / :2932-2935
          } else { EMIT(e.slot, C_QUAD); ray_nodes++; }
        } else {
          EMIT(e.slot, C_PROC); ray_nodes++;
          // which instance this procedural visit belongs to (:2171-2203 / :2951-2984 use it for the intersection table)
          const uint32_t j = ray_nodes >> 20;
          if (cnt + j < cap) rstage[cap - 1u - j] = e_inst(e);
          if (j < 0xFFFu && (ray_nodes & 0xFFFFFu) != 0xFFFFFu) ray_nodes += 1u << 20; else err |= EF_UNSUPPORTED;
        }
      }
    }
  }
#undef EMIT
#undef STAT_DEPTH
#undef PUSH_CUR
#undef STACK_FULL
#undef STK
#undef LEAF_FETCH
#undef SET_SELFROOT
#undef PUSH_OTH
#undef TAKE
#undef CUR_TID
#undef ACTIVATE
#undef LOAD_WORLD

  // ---- functional counters: one set of atomics per CTA
  atomicMax(&s_cnt[2], (max_meta >> 23) & 0xffu);
  if (err) atomicOr(&s_cnt[7], err);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* c = p.counters->v;
    const unsigned int s_nodes = s_cnt[0], m_nodes = s_cnt[1], m_lvl = s_cnt[2], s_hit = s_cnt[3], s_any = s_cnt[4], s_term = s_cnt[5], s_act = s_cnt[6], e_all = s_cnt[7];
    if (s_nodes) atomicAdd(c + CI_TOT_NODES, (unsigned long long)s_nodes);
    if (s_hit) atomicAdd(c + CI_NUM_HITS, (unsigned long long)s_hit);
    if (s_any) atomicAdd(c + CI_NUM_ANY_HITS, (unsigned long long)s_any);
    if (s_term) atomicAdd(c + CI_N_ANYHIT_RAYS, (unsigned long long)s_term);
    if (s_act - s_term) atomicAdd(c + CI_N_CLOSEST_RAYS, (unsigned long long)(s_act - s_term));
    if (s_act) atomicAdd(c + CI_RAY_COUNT, (unsigned long long)s_act);
    atomicMax(c + CI_MAX_NODES, (unsigned long long)m_nodes);
    atomicMax(c + CI_MAX_DEPTH, (unsigned long long)m_lvl);
    if (e_all) atomicOr(p.err_flags, e_all);
  }
}

template <int MODE, int STACK_N, bool EXACT, bool TNP>
int launch_mode(const TraverseParams& p, cudaStream_t st) {
  // occupancy of this instantiation, cached per device (contexts on different GPUs may live in one process)
  static int s_blocks[64], s_sm[64];
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (s_blocks[dev] == 0) {
    cudaDeviceGetAttribute(&s_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_blocks[dev], k_traverse<MODE, STACK_N, EXACT, TNP>, THREADS, 0) != cudaSuccess || s_blocks[dev] < 1) s_blocks[dev] = 4;
    // the smallest shared-memory carve-out that holds the resident CTAs; the rest of the 256 KB is L1, which the node and stack
    // lines live in (vsrt_internal.h).  VSRT_K1_CARVEOUT = percent of the maximum, -1 = the driver's default (A/B).
    { const char* e = getenv("VSRT_K1_CARVEOUT"); vsrt_min_carveout(k_traverse<MODE, STACK_N, EXACT, TNP>, s_blocks[dev], dev, e ? atoi(e) : -2); }
  }
  const int blocks_per_sm = s_blocks[dev], n_sm = s_sm[dev];
  // persistent grid (a multiple of the SM count): every resident warp keeps pulling rays until the counter runs out
  const uint64_t want = (p.n_rays + THREADS - 1) / THREADS;
  const unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)blocks_per_sm * (uint64_t)n_sm);
  if (grid == 0) return VSRT_OK;
  // p.next_ray: this launch's ray counter, zeroed by the caller (k_batch_prepare)
  if (TNP && VSRT_K1_TNODES && VSRT_K1_NODE_ENTRY && !EXACT) {
    if (!p.tv.tnodes) return VSRT_E_INVALID;      // (formation always fills the traversal copy)
    // this instantiation traverses the traversal copy: the same slots, internal nodes re-laid-out by K0, leaves and headers verbatim
    TraverseParams q = p; q.av.base = p.tv.tnodes;
    k_traverse<MODE, STACK_N, EXACT, TNP><<<grid, THREADS, 0, st>>>(q);
  } else
  k_traverse<MODE, STACK_N, EXACT, TNP><<<grid, THREADS, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

template <int STACK_N>
int launch_n(const TraverseParams& p, bool exact, bool trav_layout, cudaStream_t st) {
  if (p.mode == VSRT_MODE_TREELET) {
    if (exact) return launch_mode<VSRT_MODE_TREELET, STACK_N, true, false>(p, st);
    return trav_layout ? launch_mode<VSRT_MODE_TREELET, STACK_N, false, true>(p, st) : launch_mode<VSRT_MODE_TREELET, STACK_N, false, false>(p, st);
  }
  if (exact) return launch_mode<VSRT_MODE_DFS, STACK_N, true, false>(p, st);
  return trav_layout ? launch_mode<VSRT_MODE_DFS, STACK_N, false, true>(p, st) : launch_mode<VSRT_MODE_DFS, STACK_N, false, false>(p, st);
}

// Coherence of a batch from 256 pairs of consecutive rays spread evenly over it: *out = 1 if at least three quarters of the pairs
// point the same way to within ~2.5 degrees (cos > 0.999) -- camera rays of neighbouring pixels do, diffuse bounce rays do not
// (neighbouring origins, unrelated directions).  One warp-sized block; decides which node layout K1 reads for this batch.
__global__ void __launch_bounds__(256) k_ray_coherence(const vsrt_ray* __restrict__ rays, uint64_t n, uint32_t* __restrict__ out) {
  __shared__ unsigned int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint64_t i = n < 512 ? (uint64_t)threadIdx.x * 2u : (uint64_t)threadIdx.x * ((n - 1) / 256u);
  bool ok = false;
  if (i + 1 < n) {
    const float ax = __ldg(&rays[i].direction[0]), ay = __ldg(&rays[i].direction[1]), az = __ldg(&rays[i].direction[2]);
    const float bx = __ldg(&rays[i + 1].direction[0]), by = __ldg(&rays[i + 1].direction[1]), bz = __ldg(&rays[i + 1].direction[2]);
    const float d = ax * bx + ay * by + az * bz, na = ax * ax + ay * ay + az * az, nb = bx * bx + by * by + bz * bz;   // scheduling only: no bit-exactness contract here
    ok = d > 0.0f && d * d > 0.998f * na * nb;
  } else ok = true;
  if (ok) atomicAdd(&s_n, 1u);
  __syncthreads();
  if (threadIdx.x == 0) *out = s_n >= 192u ? 1u : 0u;
}

}  // namespace

int vsrt_launch_traverse(const TraverseParams& p, uint32_t stack_entries, bool exact, bool trav_layout, cudaStream_t st) {
  if (stack_entries <= 96) return launch_n<96>(p, exact, trav_layout, st);
  if (stack_entries <= 192) return launch_n<192>(p, exact, trav_layout, st);
  return launch_n<384>(p, exact, trav_layout, st);
}

// bytes of the global-memory traversal stack (VSRT_K1_GSTACK builds; 0 otherwise): every lane of the largest persistent grid
size_t vsrt_traverse_gstack_bytes(uint32_t stack_entries) {
#if VSRT_K1_GSTACK && VSRT_K1_NODE_ENTRY
  int dev = 0, n_sm = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const size_t sn = stack_entries <= 96 ? 96 : (stack_entries <= 192 ? 192 : 384);
  return (size_t)n_sm * 16u /* resident CTAs per SM, upper bound */ * THREADS * sn * 16u;
#else
  (void)stack_entries; return 0;
#endif
}

int vsrt_launch_ray_coherence(const vsrt_ray* rays, uint64_t n, uint32_t* out, cudaStream_t st) {
  k_ray_coherence<<<1, 256, 0, st>>>(rays, n, out);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

#if VSRT_K1_STATS
// debug build only (not part of include/vsrt.h): reads and clears the lane-state counters
extern "C" int vsrt_debug_k1_depth(unsigned long long out[128]) {
  unsigned long long z[128] = { 0 };
  if (cudaMemcpyFromSymbol(out, g_k1_depth, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_k1_depth, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
extern "C" int vsrt_debug_k1_stats(unsigned long long out[16]) {
  unsigned long long z[16] = { 0 };
  if (cudaMemcpyFromSymbol(out, g_k1_stats, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_k1_stats, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

