// traverse_wf.cu -- K1, warp-wavefront formulation of the same traversal as traverse.cu.
//
// traverse.cu gives every lane one ray and keeps its state in registers; lanes then wait for each other whenever
// their rays are in different phases (21.6 of 32 lanes in the internal-node block, 7.6 in the triangle block,
// profiles/README.md).  Here a warp owns a POOL of 64 rays whose state lives in shared memory (structure of arrays,
// 29 words per ray) and whose two LIFOs live in a per-warp global-memory stack.  Every iteration the warp takes a census of
// its pool (ballots over the state words), picks the phase with the most ready rays -- internal node, BLAS leaf,
// instance leaf, or "recycle" (write the hit record of a finished ray and start a new one in its slot) -- hands the first
// 32 ready rays to its lanes through a 32-byte selection array, and runs that ONE phase at (close to) full width.  A ray is
// still processed strictly in its own order (pop, node, pushes, pop ...), so the visit sequence is the reference's;
// regrouping only changes which lane does the work.  No CTA-level synchronisation and no atomics except the global ray
// counter: warps are independent.
//
// The phase bodies are those of traverse.cu (same device helpers, same bit-exact arithmetic).
#include "vsrt_device.cuh"
#include <algorithm>

namespace {

constexpr uint32_t INST_NONE = 0x7FFFFFu;
constexpr uint32_t RAY_DEFERRED = 0xFFFFFFFFu;
constexpr uint32_t SLOT_MASK = 0x1FFFFFFFu, SLOT_LEAF = 0x40000000u, SLOT_SELFROOT = 0x80000000u;
VS_DEV uint32_t bit_index(uint32_t one_bit) { uint32_t i; asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(one_bit)); return i; }

constexpr int WF_WARPS = 4, WF_THREADS = WF_WARPS * 32, WF_POOL = 64;
#ifndef VSRT_WF_MIN_BLOCKS
#define VSRT_WF_MIN_BLOCKS 6
#endif
enum { ST_IDLE = 0, ST_FIN = 1, ST_INT = 2, ST_INST = 3, ST_LEAF = 4 };
enum { PH_NONE = 0, PH_INT, PH_LEAF, PH_INST, PH_REC };
// per-ray state words
enum { F_ST = 0, F_R, F_ESLOT, F_EMETA, F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_TMIN, F_TMAX, F_IX, F_IY, F_IZ, F_TMULT, F_AINST,
       F_MINT, F_MINTO, F_CLEAF, F_CINST, F_CNT, F_NODES, F_ANY, F_STK, F_CTID, F_BITS, F_WTMIN, F_WTMAX, NF };
// F_BITS: bit 0 in_cur, bit 1 tid_known, bit 2 active context is non-finite, bits 8.. ray flags
constexpr uint32_t B_INCUR = 1u, B_TIDKNOWN = 2u, B_NONFINITE = 4u;

template <int MODE, bool EXACT>
__global__ void __launch_bounds__(WF_THREADS, VSRT_WF_MIN_BLOCKS) k_traverse_wf(const TraverseParams p) {
  if (p.gate && !(*reinterpret_cast<const volatile uint32_t*>(p.err_flags) & p.gate)) return;   // nothing was deferred to this pass
  const ArenaView& av = p.av;
  const uint8_t* __restrict__ base = av.base;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t inst_base = av.inst_base;   // lowest instance-leaf slot of the TLAS (K0)
  const bool only_deferred = EXACT && p.only_deferred != 0;
  const uint32_t cap = p.cap;
  const int STACK_N = (int)p.stack_n;

  __shared__ uint32_t S_[NF][WF_WARPS][WF_POOL];
  __shared__ uint8_t s_sel[WF_WARPS][32];
  __shared__ unsigned int s_cnt[8];   // 0 sum_nodes 1 max_nodes 2 max_level 3 n_hit 4 n_any 5 n_term 6 n_rays_done 7 err
  __shared__ uint32_t s_root[3];      // 0 slot word of the TLAS root entry 1 it starts in `current` 2 start treelet index
#define S(f_, s_) S_[f_][w][s_]
#define SF(f_, s_) __uint_as_float(S_[f_][w][s_])
#define SETF(f_, s_, v_) S_[f_][w][s_] = __float_as_uint(v_)
  S(F_ST, lane) = ST_IDLE; S(F_ST, lane + 32) = ST_IDLE;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    uint32_t top_root = 0, word = 0, incur = 1u, stid = VSRT_NO_TID;
    if (header_root(av, av.tlas_slot, top_root)) {
      word = top_root;
      if (MODE == VSRT_MODE_TREELET) {
        stid = root_rank(p.tv, av.tlas_slot);                               // current_treelet_root = TLAS + offset, :1707
        const uint32_t tr = __ldg(p.tv.node_tid + top_root);
        incur = ((tr & VSRT_TID_MASK) == stid) ? 1u : 0u;
        if (!incur && (tr & VSRT_TID_SELF_ROOTED)) word |= SLOT_SELFROOT;
      }
    } else word = 0xFFFFFFFFu;
    s_root[0] = word; s_root[1] = incur; s_root[2] = stid;
  }
  __syncthreads();
  uint32_t max_level = 0, err = 0;
  bool exhausted = false;
  // the warp's stack area: entry i of pool slot s at gstk[i * WF_POOL + s] (slots interleaved, like local memory interleaves lanes)
  uint2* __restrict__ gstk = p.gstack + (uint64_t)(blockIdx.x * WF_WARPS + w) * (uint64_t)STACK_N * WF_POOL;
#define STK(i_, s_) gstk[(uint64_t)(i_) * WF_POOL + (s_)]

  // ---- the ray's next entry: pops `current`, else moves the front of `other` over (:1748-1754), else the ray is finished.
  // fwd_cur_ / fwd_oth_: the entry is still in registers (just pushed), no load needed.
#define TAKE_FROM_CUR(ent_) do { eslot = (ent_).x; emeta = (ent_).y; cur_n--; bits |= B_INCUR; } while (0)
#define TAKE_FROM_OTH(ent_) do { eslot = (ent_).x; emeta = (ent_).y; oth_n--; \
      const bool selfroot_ = (eslot & SLOT_SELFROOT) != 0u; \
      if (av.tlas_delta == 0) { \
        cur_tid = eslot & SLOT_MASK; bits = (bits & ~(B_INCUR | B_TIDKNOWN)) | (selfroot_ ? B_INCUR : 0u); \
        if (!selfroot_) { cur_tid = root_rank(p.tv, eslot & SLOT_MASK); bits |= B_TIDKNOWN; } \
      } else { uint32_t s2_; bits = (bits & ~B_INCUR) | B_TIDKNOWN; \
        cur_tid = host_to_slot(av, slot_to_host(av, eslot & SLOT_MASK) - (uint64_t)av.tlas_delta, s2_) ? root_rank(p.tv, s2_) : VSRT_NO_TID; } } while (0)
#define POP_NEXT(have_c_, ent_c_, have_o_, ent_o_) do { \
      const bool fc_ = (have_c_) || cur_n > 0; \
      const bool fo_ = !fc_ && MODE == VSRT_MODE_TREELET && ((have_o_) || oth_n > 0); \
      if (fc_ || fo_) { \
        uint2 t_ = fc_ ? (ent_c_) : (ent_o_); \
        if (fc_ ? !(have_c_) : !(have_o_)) t_ = STK(fc_ ? cur_n - 1 : STACK_N - oth_n, s);   /* older entry: one load */ \
        if (fc_) TAKE_FROM_CUR(t_); else TAKE_FROM_OTH(t_); \
        const bool leaf_ = (eslot & SLOT_LEAF) != 0u; eslot &= SLOT_MASK; \
        nst = !leaf_ ? ST_INT : ((emeta & INST_NONE) == INST_NONE ? ST_INST : ST_LEAF); \
      } else nst = ST_FIN; } while (0)
#define CUR_TID() ((bits & B_TIDKNOWN) ? cur_tid : (bits |= B_TIDKNOWN, cur_tid = __ldg(p.tv.node_tid + cur_tid) & VSRT_TID_MASK))
#define EMIT(slot_, code_) do { if (cnt < cap) rstage[cnt] = ((slot_) << 3) | (uint32_t)(code_); cnt++; } while (0)
  // the ray's record for context `inst_` (INST_NONE = world) into the slot's active-ray words
#define ACTIVATE(inst_) do { const uint32_t i_ = (inst_); if (S(F_AINST, s) != i_) { \
      const vsrt_ray* rp_ = p.rays + r; Ray8 w_; \
      w_.ox = __ldg(&rp_->origin[0]); w_.oy = __ldg(&rp_->origin[1]); w_.oz = __ldg(&rp_->origin[2]); w_.tmin = SF(F_WTMIN, s); \
      w_.dx = __ldg(&rp_->direction[0]); w_.dy = __ldg(&rp_->direction[1]); w_.dz = __ldg(&rp_->direction[2]); w_.tmax = SF(F_WTMAX, s); \
      Ray8 ar_; Idir ai_; float tm_; bool nf_; \
      if (i_ == INST_NONE) { ar_ = w_; ai_ = calc_idir(w_); tm_ = 1.0f; nf_ = false; } \
      else { InstCtx c_; make_object_ray(base, inst_base + i_, w_, c_); ar_ = c_.ray; ai_ = c_.idir; tm_ = c_.tmult; nf_ = c_.exact; } \
      SETF(F_OX, s, ar_.ox); SETF(F_OY, s, ar_.oy); SETF(F_OZ, s, ar_.oz); SETF(F_DX, s, ar_.dx); SETF(F_DY, s, ar_.dy); SETF(F_DZ, s, ar_.dz); \
      SETF(F_TMIN, s, ar_.tmin); SETF(F_TMAX, s, ar_.tmax); SETF(F_IX, s, ai_.x); SETF(F_IY, s, ai_.y); SETF(F_IZ, s, ai_.z); SETF(F_TMULT, s, tm_); \
      S(F_AINST, s) = i_; bits = nf_ ? (bits | B_NONFINITE) : (bits & ~B_NONFINITE); } } while (0)

  while (true) {
    __syncwarp();
    // ================= census of the pool: lane l is the home of slots l and l + 32
    const uint32_t s0 = S(F_ST, lane), s1 = S(F_ST, lane + 32);
    const bool rec0 = s0 == ST_FIN || (s0 == ST_IDLE && !exhausted), rec1 = s1 == ST_FIN || (s1 == ST_IDLE && !exhausted);
    // how many rays are ready for each phase: one REDUX.ADD over byte counters (a pool has 64 slots)
    const uint32_t mine = (s0 == ST_INT ? 1u : 0u) + (s1 == ST_INT ? 1u : 0u) + ((s0 == ST_LEAF ? 1u : 0u) + (s1 == ST_LEAF ? 1u : 0u)) * 0x100u +
                          ((s0 == ST_INST ? 1u : 0u) + (s1 == ST_INST ? 1u : 0u)) * 0x10000u + ((rec0 ? 1u : 0u) + (rec1 ? 1u : 0u)) * 0x1000000u;
    const uint32_t tot = __reduce_add_sync(full, mine);
    const int n_int = (int)(tot & 0xffu), n_leaf = (int)((tot >> 8) & 0xffu), n_inst = (int)((tot >> 16) & 0xffu), n_rec = (int)(tot >> 24);
    if (tot == 0u) break;     // every slot idle and no ray left
    // the phase with the most ready rays (a full warp of internal nodes always wins)
    int phase = PH_INT, best = min(n_int, 32);
    if (min(n_leaf, 32) > best) { phase = PH_LEAF; best = min(n_leaf, 32); }
    if (min(n_rec, 32) > best) { phase = PH_REC; best = min(n_rec, 32); }
    if (min(n_inst, 32) > best) { phase = PH_INST; best = min(n_inst, 32); }
    const uint32_t want = phase == PH_INT ? ST_INT : phase == PH_LEAF ? ST_LEAF : ST_INST;
    const uint32_t m_lo = __ballot_sync(full, phase == PH_REC ? rec0 : s0 == want), m_hi = __ballot_sync(full, phase == PH_REC ? rec1 : s1 == want);
    // the first 32 ready slots, one per lane
    {
      const int c_lo = __popc(m_lo);
      if ((m_lo >> lane) & 1u) { const int k = __popc(m_lo & lt); if (k < 32) s_sel[w][k] = (uint8_t)lane; }
      if ((m_hi >> lane) & 1u) { const int k = c_lo + __popc(m_hi & lt); if (k < 32) s_sel[w][k] = (uint8_t)(lane + 32); }
    }
    __syncwarp();
    const bool active = lane < best;
    const uint32_t s = active ? s_sel[w][lane] : 0u;

    if (phase == PH_INT) {
      // ================= internal nodes (TLAS :1759-1875 / :2500-2599, BLAS :1954-2072 / :2687-2786)
      if (active) {
        uint32_t eslot = S(F_ESLOT, s), emeta = S(F_EMETA, s), bits = S(F_BITS, s), cnt = S(F_CNT, s), cur_tid = S(F_CTID, s);
        const uint32_t r = S(F_R, s);
        int cur_n = (int)(S(F_STK, s) & 0xffffu), oth_n = (int)(S(F_STK, s) >> 16);
        uint32_t* rstage = p.stage + (uint64_t)r * cap;
        uint32_t nst = ST_INT;
        const Node64 n = load_node(base, eslot);
        const uint32_t inst = emeta & INST_NONE;
        EMIT(eslot, inst == INST_NONE ? C_INTERNAL_TLAS : C_INTERNAL_BLAS);
        S(F_NODES, s) = S(F_NODES, s) + 1u;
        ACTIVATE(inst);
        if (!EXACT && (bits & B_NONFINITE)) { p.counts[r] = RAY_DEFERRED; err |= EF_NEED_EXACT; nst = ST_IDLE; }   // degenerate instance transform
        else {
          Ray8 ar; Idir ai;
          ar.ox = SF(F_OX, s); ar.oy = SF(F_OY, s); ar.oz = SF(F_OZ, s); ar.dx = ar.dy = ar.dz = 0.0f; ar.tmin = SF(F_TMIN, s); ar.tmax = SF(F_TMAX, s);
          ai.x = SF(F_IX, s); ai.y = SF(F_IY, s); ai.z = SF(F_IZ, s);
          const uint32_t mask = test_children<EXACT>(n, ar, ai, fmul(SF(F_MINT, s), SF(F_TMULT, s)), p.magic16);   // cull: :1791 / :1989
          const uint32_t lo4 = __byte_perm(n.w[5], n.w[6], 0x5432), hi2 = n.w[6] >> 16;
          const uint32_t pre4 = (lo4 & 0x03030303u) * 0x01010101u;
          const uint32_t t4 = pre4 >> 24, xlo = pre4 << 8, xhi = t4 | ((t4 + (hi2 & 3u)) << 8);
          const uint32_t child0 = eslot + (uint32_t)node_child_offset(n);
          const uint32_t level = (emeta >> 23) & 0xffu, clevel = level < 255u ? level + 1u : 255u;
          if (mask && clevel > max_level) max_level = clevel;
          const uint32_t cmeta = (clevel << 23) | inst;
          uint2 lc = make_uint2(0u, 0u), lo = make_uint2(0u, 0u); bool hc = false, ho = false;
          if (MODE == VSRT_MODE_TREELET) {
            uint32_t mc = node_byte(n, 17);
            if (!(bits & B_INCUR)) {
              mc = 0;
              const uint32_t ct = CUR_TID();
              for (uint32_t m = mask; m; m &= m - 1u) {
                const uint32_t sel = 0x7770u + bit_index(m & (0u - m));
                if ((__ldg(p.tv.node_tid + child0 + __byte_perm(xlo, xhi, sel)) & VSRT_TID_MASK) == ct) mc |= m & (0u - m);
              }
            }
            const uint32_t mcur = mask & mc;
            if (cur_n + oth_n + __popc(mask) > STACK_N) err |= EF_STACK;
            else {
              int po = STACK_N - 1 - oth_n;
              for (uint32_t m = mask; m; ) {
                const uint32_t bit = m & (0u - m); m ^= bit;
                const uint32_t sel = 0x7770u + bit_index(bit);
                const uint2 c = make_uint2((child0 + __byte_perm(xlo, xhi, sel)) | ((__byte_perm(lo4, hi2, sel) << 24) & 0xC0000000u), cmeta);
                const bool ic = (mcur & bit) != 0u;
                STK(ic ? cur_n : po, s) = c;
                if (ic) { lc = c; hc = true; cur_n++; } else { lo = c; ho = true; po--; }
              }
              oth_n = STACK_N - 1 - po;
            }
            ho = ho && cur_n == 0;
          } else {
            // the first hit internal child is followed at once (:2573); every other hit child is pushed in slot order
            if (cur_n + __popc(mask) > STACK_N) err |= EF_STACK;
            else {
              for (uint32_t m = mask; m; ) {
                const uint32_t bit = m & (0u - m); m ^= bit;
                const uint32_t sel = 0x7770u + bit_index(bit);
                const uint32_t fl = (__byte_perm(lo4, hi2, sel) << 24) & 0xC0000000u;
                const uint2 c = make_uint2((child0 + __byte_perm(xlo, xhi, sel)) | fl, cmeta);
                if (!(fl & SLOT_LEAF) && !hc) { lc = c; hc = true; cur_n++; }      // "popped" right below without ever being stored
                else { STK(hc ? cur_n - 1 : cur_n, s) = c; cur_n++; }
              }
            }
          }
          POP_NEXT(hc, lc, ho, lo);
        }
        S(F_ESLOT, s) = eslot; S(F_EMETA, s) = emeta; S(F_BITS, s) = bits; S(F_CNT, s) = cnt; S(F_CTID, s) = cur_tid;
        S(F_STK, s) = (uint32_t)cur_n | ((uint32_t)oth_n << 16);
        S(F_ST, s) = nst;
      }
    } else if (phase == PH_LEAF) {
      // ================= BLAS leaves (:2073-2204 / :2789-2985)
      if (active) {
        uint32_t eslot = S(F_ESLOT, s), emeta = S(F_EMETA, s), bits = S(F_BITS, s), cnt = S(F_CNT, s), cur_tid = S(F_CTID, s);
        const uint32_t r = S(F_R, s), flags = bits >> 8;
        int cur_n = (int)(S(F_STK, s) & 0xffffu), oth_n = (int)(S(F_STK, s) >> 16);
        uint32_t* rstage = p.stage + (uint64_t)r * cap;
        uint32_t nst = ST_LEAF;
        const Node64 q = load_node_now(base, eslot);
        EMIT(eslot, C_DESC);
        if (((q.w[1] >> 29) & 1u) == 0u) {
          ACTIVATE(emeta & INST_NONE);
          Ray8 ar;
          ar.ox = SF(F_OX, s); ar.oy = SF(F_OY, s); ar.oz = SF(F_OZ, s); ar.dx = SF(F_DX, s); ar.dy = SF(F_DY, s); ar.dz = SF(F_DZ, s);
          ar.tmin = SF(F_TMIN, s); ar.tmax = SF(F_TMAX, s);
          const float tmult = SF(F_TMULT, s), w_tmin = SF(F_WTMIN, s), w_tmax = SF(F_WTMAX, s);
          float min_thit = SF(F_MINT, s);
          float thit = 0.0f;
          const bool hit = ray_tri(q, ar, thit);
          const float tw = !hit ? 0.0f : (tmult == 1.0f ? thit : fdiv(thit, tmult));
          bool acc = hit && w_tmin <= tw && tw <= w_tmax;                         // :2843
          if (MODE == VSRT_MODE_TREELET) acc = acc && tw < min_thit;              // :2127
          if (acc) {
            if (MODE == VSRT_MODE_TREELET) min_thit = tw;
            else {
              const bool opaque = (flags & VSRT_RAY_FLAG_OPAQUE) != 0;            // skipAnyHitShader, :2413
              if (opaque && tw < min_thit) min_thit = tw;                         // :2850
              if (!opaque) S(F_ANY, s) = S(F_ANY, s) + 1u;                        // :2869-2929
            }
            SETF(F_MINT, s, min_thit); SETF(F_MINTO, s, thit); S(F_CLEAF, s) = eslot; S(F_CINST, s) = emeta & INST_NONE;
            EMIT(eslot, C_QUAD_HIT);
            if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) { cur_n = 0; oth_n = 0; }   // :2151-2155 / :2932-2935
          } else EMIT(eslot, C_QUAD);
          S(F_NODES, s) = S(F_NODES, s) + 1u;
        } else {
          EMIT(eslot, C_PROC);
          // which instance this procedural visit belongs to: kept at the end of the staging segment (see traverse.cu)
          const uint32_t nn = S(F_NODES, s) + 1u, j = nn >> 20;
          if (cnt + j < cap) rstage[cap - 1u - j] = emeta & INST_NONE;
          if (j < 0xFFFu && (nn & 0xFFFFFu) != 0xFFFFFu) S(F_NODES, s) = nn + (1u << 20); else { S(F_NODES, s) = nn; err |= EF_UNSUPPORTED; }
        }
        { const uint2 z = make_uint2(0u, 0u); POP_NEXT(false, z, false, z); }
        S(F_ESLOT, s) = eslot; S(F_EMETA, s) = emeta; S(F_BITS, s) = bits; S(F_CNT, s) = cnt; S(F_CTID, s) = cur_tid;
        S(F_STK, s) = (uint32_t)cur_n | ((uint32_t)oth_n << 16);
        S(F_ST, s) = nst;
      }
    } else if (phase == PH_INST) {
      // ================= instance leaves (:1876-1953 / :2602-2677)
      if (active) {
        uint32_t eslot = S(F_ESLOT, s), emeta = S(F_EMETA, s), bits = S(F_BITS, s), cnt = S(F_CNT, s), cur_tid = S(F_CTID, s);
        const uint32_t r = S(F_R, s);
        int cur_n = (int)(S(F_STK, s) & 0xffffu), oth_n = (int)(S(F_STK, s) >> 16);
        uint32_t* rstage = p.stage + (uint64_t)r * cap;
        uint32_t nst = ST_INST;
        EMIT(eslot, C_INSTANCE);
        S(F_NODES, s) = S(F_NODES, s) + 1u;
        uint32_t hdr = 0, broot = 0;
        const uint32_t iref = eslot - inst_base;
        bool dead = false;
        if (!instance_blas_header(av, eslot, hdr) || !header_root(av, hdr, broot)) { err |= EF_BAD_BVH; dead = true; }
        else if (eslot < inst_base || iref >= INST_NONE) { err |= EF_BAD_BVH; dead = true; }   // cannot happen for an arena K0 accepted
        else {
          EMIT(hdr, C_STRUCT);                                                     // BLAS header record, :1913 / :2645
          uint2 c = make_uint2(broot, (((emeta >> 23) & 0xffu) << 23) | iref);     // BLAS root inherits the leaf's level (:1944)
          if (MODE == VSRT_MODE_DFS) { if (cur_n < STACK_N) { STK(cur_n, s) = c; cur_n++; } else err |= EF_STACK; }
          else {
            const uint32_t tb = __ldg(p.tv.node_tid + broot);
            if (cur_n + oth_n >= STACK_N) err |= EF_STACK;
            else if ((tb & VSRT_TID_MASK) == CUR_TID()) { STK(cur_n, s) = c; cur_n++; }
            else { c.x |= (tb & VSRT_TID_SELF_ROOTED) ? SLOT_SELFROOT : 0u; oth_n++; STK(STACK_N - oth_n, s) = c; }
          }
        }
        if (dead) nst = ST_FIN;
        else { const uint2 z = make_uint2(0u, 0u); POP_NEXT(false, z, false, z); }
        S(F_ESLOT, s) = eslot; S(F_EMETA, s) = emeta; S(F_BITS, s) = bits; S(F_CNT, s) = cnt; S(F_CTID, s) = cur_tid;
        S(F_STK, s) = (uint32_t)cur_n | ((uint32_t)oth_n << 16);
        S(F_ST, s) = nst;
      }
    } else {
      // ================= recycle: hit record of a finished ray (:2211-2245 / :2990-3033), then a new ray in the slot
      const bool was_fin = active && S(F_ST, s) == ST_FIN;
      if (was_fin) {
        const uint32_t r = S(F_R, s), cnt = S(F_CNT, s), nproc = S(F_NODES, s) >> 20, ray_nodes = S(F_NODES, s) & 0xFFFFFu, ray_any = S(F_ANY, s), flags = S(F_BITS, s) >> 8;
        const float min_thit = SF(F_MINT, s), min_thit_object = SF(F_MINTO, s), w_tmax = SF(F_WTMAX, s);
        if (cnt + nproc > cap) err |= EF_TRACE_CAP;
        p.counts[r] = cnt; p.nproc[r] = nproc;
        vsrt_hit h;
        h.hit_geometry = 0; h.world_min_thit = 0.0f; h.primitive_index = 0; h.geometry_index = 0; h.instance_index = 0;
        h.barycentric[0] = h.barycentric[1] = h.barycentric[2] = 0.0f;
        h.intersection_point[0] = h.intersection_point[1] = h.intersection_point[2] = 0.0f;
        h.n_all_hits = ray_any; h.instance_leaf_address = 0;
        if (min_thit < w_tmax) {
          atomicAdd(&s_cnt[3], 1u);
          const vsrt_ray* rp = p.rays + r;
          Ray8 wr;
          wr.ox = __ldg(&rp->origin[0]); wr.oy = __ldg(&rp->origin[1]); wr.oz = __ldg(&rp->origin[2]); wr.tmin = SF(F_WTMIN, s);
          wr.dx = __ldg(&rp->direction[0]); wr.dy = __ldg(&rp->direction[1]); wr.dz = __ldg(&rp->direction[2]); wr.tmax = w_tmax;
          const uint32_t closest_leaf = S(F_CLEAF, s), ci = inst_base + S(F_CINST, s);
          const Node64 q = load_node(base, closest_leaf);
          InstCtx c; make_object_ray(base, ci, wr, c);
          h.hit_geometry = 1; h.world_min_thit = min_thit;
          h.geometry_index = q.w[1] & 0x0fffffffu; h.primitive_index = q.w[2];
          h.instance_index = __ldg(reinterpret_cast<const uint32_t*>(base + (uint64_t)ci * 64u + 72u));
          h.intersection_point[0] = fadd(wr.ox, fmul(wr.dx, min_thit));
          h.intersection_point[1] = fadd(wr.oy, fmul(wr.dy, min_thit));
          h.intersection_point[2] = fadd(wr.oz, fmul(wr.dz, min_thit));
          barycentric(q, fadd(c.ray.ox, fmul(c.ray.dx, min_thit_object)), fadd(c.ray.oy, fmul(c.ray.dy, min_thit_object)),
                      fadd(c.ray.oz, fmul(c.ray.dz, min_thit_object)), h.barycentric);
          h.instance_leaf_address = slot_to_host(av, ci);
        }
        p.hits[r] = h;
        atomicAdd(&s_cnt[0], ray_nodes); atomicMax(&s_cnt[1], ray_nodes);
        if (ray_any) atomicAdd(&s_cnt[4], ray_any);
        if (flags & VSRT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) atomicAdd(&s_cnt[5], 1u);
        atomicAdd(&s_cnt[6], 1u);
        S(F_ST, s) = ST_IDLE;
      }
      if (!exhausted) {
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(p.next_ray, (unsigned long long)best);
        b0 = __shfl_sync(full, b0, 0);
        if (b0 + (unsigned long long)best >= p.n_rays) exhausted = true;
        const uint64_t nr = b0 + (uint64_t)lane;
        if (active && nr < p.n_rays && (!only_deferred || p.counts[nr] == RAY_DEFERRED)) {
          // ---- start ray nr (:1650-1741 / :2411-2484)
          const uint32_t r = (uint32_t)nr;
          uint32_t* rstage = p.stage + nr * cap;
          const vsrt_ray* rp = p.rays + r;
          Ray8 wr;
          wr.ox = __ldg(&rp->origin[0]); wr.oy = __ldg(&rp->origin[1]); wr.oz = __ldg(&rp->origin[2]); wr.tmin = __ldg(&rp->tmin);
          wr.dx = __ldg(&rp->direction[0]); wr.dy = __ldg(&rp->direction[1]); wr.dz = __ldg(&rp->direction[2]); wr.tmax = __ldg(&rp->tmax);
          const uint32_t flags = __ldg(&rp->ray_flags);
          if (!EXACT && ray_needs_exact(wr)) { p.counts[r] = RAY_DEFERRED; err |= EF_NEED_EXACT; }   // left to the EXACT pass
          else {
            const Idir wi = calc_idir(wr);
            uint32_t cnt = 0, eslot = 0, emeta = 0, cur_tid = s_root[2], bits = B_TIDKNOWN | ((flags & 0xffffffu) << 8), nst = ST_FIN;
            int cur_n = 0, oth_n = 0;
            S(F_R, s) = r;
            SETF(F_OX, s, wr.ox); SETF(F_OY, s, wr.oy); SETF(F_OZ, s, wr.oz); SETF(F_DX, s, wr.dx); SETF(F_DY, s, wr.dy); SETF(F_DZ, s, wr.dz);
            SETF(F_TMIN, s, wr.tmin); SETF(F_TMAX, s, wr.tmax); SETF(F_IX, s, wi.x); SETF(F_IY, s, wi.y); SETF(F_IZ, s, wi.z); SETF(F_TMULT, s, 1.0f);
            S(F_AINST, s) = INST_NONE; SETF(F_WTMIN, s, wr.tmin); SETF(F_WTMAX, s, wr.tmax);
            SETF(F_MINT, s, wr.tmax); SETF(F_MINTO, s, 0.0f); S(F_CLEAF, s) = 0u; S(F_CINST, s) = INST_NONE;   // :1671
            S(F_NODES, s) = 0u; S(F_ANY, s) = 0u;
            EMIT(av.tlas_slot, C_STRUCT);                                                // :1685 / :2447
            if (s_root[0] == 0xFFFFFFFFu) err |= EF_BAD_BVH;
            else {
              // scene box of the TLAS header (:1712-1741 / :2472-2484)
              const float* hb = reinterpret_cast<const float*>(base + (uint64_t)av.tlas_slot * 64u + 8u);
              float th;
              if (ray_box(__ldg(hb), __ldg(hb + 1), __ldg(hb + 2), __ldg(hb + 3), __ldg(hb + 4), __ldg(hb + 5), wi, wr, th)) {
                const uint2 c = make_uint2(s_root[0], (1u << 23) | INST_NONE);
                const bool to_cur = MODE != VSRT_MODE_TREELET || s_root[1] != 0u;
                if (to_cur) cur_n = 1; else oth_n = 1;
                if (max_level < 1) max_level = 1;
                nst = ST_INT;
                POP_NEXT(to_cur, c, !to_cur, c);
              }
            }
            S(F_ESLOT, s) = eslot; S(F_EMETA, s) = emeta; S(F_BITS, s) = bits; S(F_CNT, s) = cnt; S(F_CTID, s) = cur_tid;
            S(F_STK, s) = (uint32_t)cur_n | ((uint32_t)oth_n << 16);
            S(F_ST, s) = nst;
          }
        }
      }
    }
  }
#undef S
#undef SF
#undef SETF
#undef STK
#undef TAKE_FROM_CUR
#undef TAKE_FROM_OTH
#undef POP_NEXT
#undef CUR_TID
#undef EMIT
#undef ACTIVATE

  // ---- functional counters: one set of atomics per CTA
  atomicMax(&s_cnt[2], max_level);
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

template <int MODE, bool EXACT>
int launch_wf(const TraverseParams& p, unsigned grid, cudaStream_t st) {
  k_traverse_wf<MODE, EXACT><<<grid, WF_THREADS, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}

}  // namespace

// persistent grid: SMs x resident CTAs, never more warps than there are pools of rays
unsigned vsrt_wf_grid(uint64_t n_rays) {
  static int n_sm = 0, per_sm = 0;
  if (n_sm == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_wf<VSRT_MODE_TREELET, false>, WF_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  }
  const uint64_t want = (n_rays + (uint64_t)WF_WARPS * WF_POOL - 1) / ((uint64_t)WF_WARPS * WF_POOL);
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)per_sm * (uint64_t)n_sm));
}
size_t vsrt_wf_stack_bytes(unsigned grid, uint32_t stack_entries) { return (size_t)grid * WF_WARPS * WF_POOL * (size_t)stack_entries * sizeof(uint2); }

int vsrt_launch_traverse_wf(const TraverseParams& p, unsigned grid, bool exact, cudaStream_t st) {
  if (p.n_rays == 0) return VSRT_OK;
  if (cudaMemsetAsync(p.next_ray, 0, sizeof(unsigned long long), st) != cudaSuccess) return VSRT_E_CUDA;
  if (p.mode == VSRT_MODE_TREELET) return exact ? launch_wf<VSRT_MODE_TREELET, true>(p, grid, st) : launch_wf<VSRT_MODE_TREELET, false>(p, grid, st);
  return exact ? launch_wf<VSRT_MODE_DFS, true>(p, grid, st) : launch_wf<VSRT_MODE_DFS, false>(p, grid, st);
}
