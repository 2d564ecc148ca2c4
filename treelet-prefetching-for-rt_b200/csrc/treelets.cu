// treelets.cu -- K0: treelet formation, the GPU restatement of VulkanRayTracing::createTreelets
// (vulkan_ray_tracing.cc:823-1470) + buildNodeToRootMap (:475-489).
//
// The reference is a sequential top-down greedy BFS: a FIFO of candidate nodes is drained while the FRONT
// candidate still fits the byte budget (:1117); when it does not, the treelet closes and EVERY queued candidate
// becomes the root of a later treelet (:1167-1171).  A treelet's content is therefore a pure function of
// (root, budget), and all roots discovered by one generation of treelets are independent: the kernel below forms
// one generation ("wave") per launch, one thread per root, and the host loops over waves until no new root
// appears.  Results are identical to the sequential algorithm because the final tables are keyed by address:
//   * roots ascending by address  -> rank == treelet_addr_to_metadata_idx (:1311-1334)
//   * node lists in BFS order, de-duplicated first-occurrence-wins (:1312-1330)
//   * node -> root: "later (higher-address) roots overwrite" (:479-486)  == atomicMax over ranks.
// The FIFO is never materialised: the processed-node list IS the queue prefix, so the queue front is "the next
// unconsumed child of list[p]" and only a cursor (p, child index) is kept per thread.
// The walk also validates what the reference asserts on (:926,:973,:2108 and every remaining_bytes >= 0).
#include "vsrt_device.cuh"
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>

namespace {

struct FormState {
  uint32_t* claimed;           // root bitmap
  uint2* roots;                // worklist: (slot, kind)
  unsigned int* n_roots;
  unsigned long long* total_bvh;
  uint32_t* err;
  uint32_t* inst_range;        // [0] lowest, [1] highest instance-leaf slot met (K1 names an instance by its slot relative to [0])
  uint64_t* store;             // compact store of finished node lists
  unsigned long long* cursor;  // entries of the store handed out so far
  unsigned long long store_cap;
  unsigned long long* r_off;   // [root] where its list starts in the store
};

VS_DEV uint64_t mk_entry(uint32_t slot, uint32_t kind) { return (uint64_t)slot | ((uint64_t)kind << 32); }

struct Walker {
  const ArenaView& av; uint64_t* list; uint32_t n; int remaining; uint32_t err; uint32_t n_inst; unsigned long long bytes; uint32_t inst_lo, inst_hi;
  VS_DEV Walker(const ArenaView& a, uint64_t* l, int budget) : av(a), list(l), n(0), remaining(budget), err(0), n_inst(0), bytes(0), inst_lo(0xFFFFFFFFu), inst_hi(0) {}
  VS_DEV void charge(int b) { remaining -= b; bytes += (unsigned)b; if (remaining < 0) err |= EF_BUDGET; }   // assert(remaining_bytes >= 0)
  // "process" one popped candidate: append to the node list and charge its bytes (:886-1114)
  VS_DEV void process(uint32_t slot, uint32_t kind) {
    if (slot >= av.n_slots) { err |= EF_BAD_BVH; return; }
    if (kind == K_TLAS_INTERNAL || kind == K_BLAS_INTERNAL) {
      charge(64); list[n++] = mk_entry(slot, kind);
      {   // a non-finite node origin lets a NaN reach the slab test: traversal must then keep the ternary MIN/MAX
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(av.base + (uint64_t)slot * 64u));
        if (!finite3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z))) err |= EF_NONFINITE;
        // so does a present child whose quantised lower bound exceeds its upper bound on some axis: the fast slab test takes
        // the near / far plane from the sign of the ray direction, which equals min / max only for ordered bounds
        const uint4* np = reinterpret_cast<const uint4*>(av.base + (uint64_t)slot * 64u);
        const uint4 b = __ldg(np + 1), c = __ldg(np + 2), d = __ldg(np + 3);
        const uint32_t w[16] = { 0, 0, 0, 0, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w };
        // and a bound exponent below -118 (scale 2^(e-8) is a denormal): the fast path builds the scale from the byte alone
#pragma unroll
        for (int ax = 0; ax < 3; ax++) if ((int)(int8_t)((w[(18 + ax) >> 2] >> (((18 + ax) & 3) * 8)) & 0xffu) < -118) err |= EF_NONFINITE;
#pragma unroll
        for (int i = 0; i < 6; i++) {
          if (((w[(22 + i) >> 2] >> (((22 + i) & 3) * 8)) & 3u) == 0u) continue;
#pragma unroll
          for (int ax = 0; ax < 3; ax++) {
            const int lo = 28 + 12 * ax + i, hi = lo + 6;
            if (((w[lo >> 2] >> ((lo & 3) * 8)) & 0xffu) > ((w[hi >> 2] >> ((hi & 3) * 8)) & 0xffu)) err |= EF_NONFINITE;
          }
        }
      }
      if (kind == K_TLAS_INTERNAL) {   // assert(node.ChildType[i] == NODE_TYPE_INSTANCE) for TLAS leaves, :926
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(av.base + (uint64_t)slot * 64u) + 1);
        const uint64_t info6 = ((uint64_t)b.z << 16) | (b.y >> 16);
#pragma unroll
        for (int i = 0; i < 6; i++) { const uint32_t t = (uint32_t)(info6 >> (8 * i)) & 0x3fu; if ((t & 3u) && (t >> 2) > 1u) err |= EF_BAD_BVH; }
      }
    } else if (kind == K_INSTANCE) {
      if (slot + 1 >= av.n_slots) { err |= EF_BAD_BVH; return; }
      charge(128); list[n++] = mk_entry(slot, K_INSTANCE); n_inst++;
      inst_lo = min(inst_lo, slot); inst_hi = max(inst_hi, slot);
      uint32_t hdr = 0; int64_t d;
      if (!instance_blas_header(av, slot, hdr)) { err |= EF_BAD_BVH; return; }
      if (!blas_delta_of(av, hdr, d)) { err |= EF_UNKNOWN_AS; return; }                                    // assert :973
      charge(64); list[n++] = mk_entry(hdr, K_BLAS_HEADER);                                               // isBlasRoot entry, :975
    } else {   // BLAS leaf: 64 bytes whether quad or procedural (:1100,:1109)
      list[n++] = mk_entry(slot, K_BLAS_LEAF); charge(64);
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(av.base + (uint64_t)slot * 64u));
      if (((a.y >> 29) & 1u) == 0u && (a.w & 0x1ffffu) != 0u) err |= EF_BAD_BVH;                          // PrimitiveIndex1Delta, :2108
    }
  }
};

// One thread per root of the current wave.
// The walk needs room for budget / 64 + 2 entries per root, most roots need a handful: the lists are built in a scratch area
// that only has to hold the roots of one launch and are then copied, at their real length, into a compact store (the space is
// handed out by an atomic cursor; a launch whose lists do not fit raises EF_TRACE_CAP and is repeated after the host has grown
// the store -- it is idempotent up to that point: roots are claimed once, totals are added only with the copy).
__global__ void __launch_bounds__(128) k_form_wave(const ArenaView av, const FormState fs, uint32_t begin, uint32_t count, int budget,
                                                   uint32_t cap, uint64_t* __restrict__ pool, uint32_t* __restrict__ r_count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < count;                                   // (no early return: the whole warp allocates store space together below)
  const uint2 root = active ? fs.roots[begin + t] : make_uint2(0u, 0xFFFFFFFFu);
  uint64_t* list = pool + (uint64_t)(active ? t : 0u) * cap;
  Walker w(av, list, budget);
  uint32_t p = 0;
  if (!active) { }
  else if (root.y == K_TLAS_HEADER) {                                   // :846-862
    w.charge(64); list[w.n++] = mk_entry(root.x, K_TLAS_HEADER);
    uint32_t rs = 0;
    if (!header_root(av, root.x, rs)) w.err |= EF_BAD_BVH; else w.process(rs, K_TLAS_INTERNAL);
    p = 1;
  } else w.process(root.x, root.y);                                // a pending root is processed without a fit check (:1197)

  // queue cursor: children of list[p] from child index ci on
  uint32_t ci = 0, child = 0; uint64_t info6 = 0; bool loaded = false; bool closing = false;
  while (active && p < w.n && !(w.err & (EF_BAD_BVH | EF_UNKNOWN_AS))) {
    const uint64_t ent = list[p];
    const uint32_t eslot = (uint32_t)ent, ekind = (uint32_t)(ent >> 32);
    uint32_t cslot = 0, ckind = 0, csz = 0; bool found = false;
    if (ekind == K_TLAS_INTERNAL || ekind == K_BLAS_INTERNAL) {
      if (!loaded) {
        const uint4* np = reinterpret_cast<const uint4*>(av.base + (uint64_t)eslot * 64u);
        const uint4 a = __ldg(np), b = __ldg(np + 1);
        child = eslot + a.w;                                        // ChildOffset in 64-byte units, relative to the node (:919)
        info6 = ((uint64_t)b.z << 16) | (b.y >> 16);                // bytes 22..27
        ci = 0; loaded = true;
      }
      while (ci < 6) {
        const uint32_t t6 = (uint32_t)(info6 >> (8 * ci)) & 0x3fu;
        if (t6 & 3u) { cslot = child; csz = t6 & 3u; const uint32_t ty = t6 >> 2;
          ckind = (ekind == K_TLAS_INTERNAL) ? (ty == 0 ? K_TLAS_INTERNAL : K_INSTANCE) : (ty == 0 ? K_BLAS_INTERNAL : K_BLAS_LEAF);
          found = true; break; }
        ci++;
      }
    } else if (ekind == K_INSTANCE && ci == 0) {
      // its single queued child is the BLAS root internal node (:987-990); the header entry sits at list[p+1]
      const uint32_t hdr = (uint32_t)list[p + 1];
      uint32_t rs = 0;
      if (header_root(av, hdr, rs)) { cslot = rs; ckind = K_BLAS_INTERNAL; csz = 0; found = true; } else { w.err |= EF_BAD_BVH; break; }
    }
    if (!found) { p++; ci = 0; loaded = false; continue; }
    if (!closing && w.remaining - (ckind == K_INSTANCE ? 192 : 64) >= 0) {      // front fits (:1117)
      ci++; child += csz;
      w.process(cslot, ckind);
    } else {
      // treelet closed: every queued candidate becomes a future root (:1167-1171)
      closing = true;
      ci++; child += csz;
      if (cslot >= av.n_slots) { w.err |= EF_BAD_BVH; break; }
      const uint32_t bit = 1u << (cslot & 31);
      const uint32_t old = atomicOr(fs.claimed + (cslot >> 5), bit);
      if (!(old & bit)) { const unsigned int idx = atomicAdd(fs.n_roots, 1u); fs.roots[idx] = make_uint2(cslot, ckind); }
    }
  }
  // de-duplicate, first occurrence wins (:1312-1330).  A node can only repeat inside one treelet when two instance
  // leaves of that treelet reference the same BLAS.
  uint32_t n = w.n;
  if (w.n_inst >= 2) {
    uint32_t o = 0;
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t s = (uint32_t)list[i]; bool dup = false;
      for (uint32_t j = 0; j < o; j++) if ((uint32_t)list[j] == s) { dup = true; break; }
      if (!dup) list[o++] = list[i];
    }
    n = o;
  }
  if (w.err) atomicOr(fs.err, w.err);
  const bool keep = active && !(w.err & (EF_BAD_BVH | EF_UNKNOWN_AS | EF_BUDGET));
  if (!keep) n = 0;
  // space in the compact store: one atomic per WARP (a cursor bumped once per root serialises in L2 -- 0.6 M same-address atomics
  // cost 3 ms of a 5 ms formation), the lanes take consecutive pieces of the warp's range
  __syncwarp();
  const int lane = threadIdx.x & 31;
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
  const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long wbase = 0;
  if (lane == 31 && warp_total) wbase = atomicAdd(fs.cursor, (unsigned long long)warp_total);
  wbase = __shfl_sync(0xffffffffu, wbase, 31);
  if (!keep) return;
  const unsigned long long off = wbase + (incl - n);
  if (off + n > fs.store_cap) { atomicOr(fs.err, (uint32_t)EF_TRACE_CAP); return; }
  for (uint32_t i = 0; i < n; i++) fs.store[off + i] = list[i];
  fs.r_off[begin + t] = off;
  r_count[begin + t] = n;
  atomicAdd(fs.total_bvh, w.bytes);
  if (w.n_inst) { atomicMin(fs.inst_range, w.inst_lo); atomicMax(fs.inst_range + 1, w.inst_hi); }
}

__global__ void k_popc(const uint32_t* __restrict__ bits, uint32_t nw, uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < nw) cnt[i] = __popc(bits[i]);
}
__global__ void k_narrow(const unsigned long long* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = (uint32_t)in[i];
}
// rank of every discovered root + scatter of (slot, count) into rank order
// ... and the owner of every bucket of K3's hot table: the treelet discovered first (lowest i: roots are appended generation by
// generation, starting at the TLAS) among those whose index falls into the bucket
__global__ void k_rank(const uint2* __restrict__ roots, const uint32_t* __restrict__ r_count, uint32_t n, const uint32_t* __restrict__ bits,
                       const uint32_t* __restrict__ prefix, uint32_t* __restrict__ r_rank, uint32_t* __restrict__ tl_root, uint32_t* __restrict__ tl_count,
                       unsigned long long* __restrict__ hot64) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const uint32_t s = roots[i].x, w = bits[s >> 5], b = 1u << (s & 31);
  const uint32_t rk = prefix[s >> 5] + __popc(w & (b - 1));
  r_rank[i] = rk; tl_root[rk] = s; tl_count[rk] = r_count[i];
  atomicMin(hot64 + (rk & (VSRT_HOT_N - 1u)), ((unsigned long long)i << 32) | rk);
}
// copy each root's list into the rank-ordered CSR and fold the node -> highest-root map (rank + 1, 0 = unmapped)
__global__ void k_gather(const uint64_t* __restrict__ store, const unsigned long long* __restrict__ r_off, const uint32_t* __restrict__ r_count, const uint32_t* __restrict__ r_rank, uint32_t n,
                         const unsigned long long* __restrict__ tl_off, uint64_t* __restrict__ tl_node, uint32_t* __restrict__ node_tid1,
                         unsigned long long* __restrict__ n_mapped) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const uint64_t* src = store + r_off[i]; const uint32_t c = r_count[i], rk = r_rank[i];
  uint64_t* dst = tl_node + tl_off[rk];
  uint32_t fresh = 0;
  for (uint32_t k = 0; k < c; k++) {
    const uint64_t e = src[k]; dst[k] = e;
    const uint32_t old = atomicMax(node_tid1 + (uint32_t)e, rk + 1u);
    if (old == 0u) fresh++;
  }
  if (fresh) atomicAdd(n_mapped, (unsigned long long)fresh);
}
// node_tid[slot] = treelet index | VSRT_TID_SELF_ROOTED when the slot is itself the root of that treelet (the common
// case for a treelet root; it differs only when a shared BLAS puts a root inside a higher-addressed treelet too)
__global__ void k_fix_tid(uint32_t* __restrict__ node_tid, uint32_t n, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ prefix) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = node_tid[i];
  if (!v) { node_tid[i] = VSRT_NO_TID; return; }
  uint32_t t = v - 1u;
  const uint32_t w = bits[i >> 5], b = 1u << (i & 31);
  if ((w & b) && prefix[i >> 5] + __popc(w & (b - 1)) == t) t |= VSRT_TID_SELF_ROOTED;
  node_tid[i] = t;
}
// Per internal node, written into spare bits of the context's private arena copy so that K1 gets them with the node's own
// 64 bytes instead of gathering node_tid per child:
//   pad byte +17 ("one unused byte" of GEN_RT_BVH_INTERNAL_NODE_unpack, util.h:146), bit i: child slot i exists and is
//     mapped to the SAME treelet as the node;
//   bit 7 of child-info byte 22+i (the reference reads these bytes & 0x3f, util.h:160): child i is the root of the treelet
//     it is mapped to (VSRT_TID_SELF_ROOTED of node_tid[child]; also set for an unmapped child, like the flag itself);
//   bit 6 of the same byte: child i is a leaf (ChildType != 0).  Bits 7/6 are bits 31/30 of a traversal-stack entry.
// One thread per list entry; a node listed by several treelets (shared BLAS) gets the same value from each.
//
// The same thread also writes the node in K1's TRAVERSAL LAYOUT into `tnodes`, a second copy of the arena (same slots; leaves, instance
// leaves and headers verbatim): everything the hot kernel derives from an internal node per visit with byte shuffles is laid down once --
//   w0..2 origin | w3 slot of the first child (absolute) | w4 per-child byte "offset | leaf << 6 | self-rooted << 7", children 0..3
//   w5 the same bytes of children 4,5 | present mask << 16 | leaf mask << 24
//   w6 bytes 0..2: (exponent + 119) & 255 per axis, i.e. bits 23..30 of the scale 2^(e-8) as a float; byte 3: the same-treelet mask
//   w7 + 3 * axis: lower bounds of children 0..3 | upper bounds of children 0..3 | lower 4, lower 5, upper 4, upper 5
// so that near / far planes are whole words chosen by the sign of the ray direction.
__global__ void k_child_mask(uint8_t* __restrict__ arena, uint32_t n_slots, const uint64_t* __restrict__ tl_node, unsigned long long n_entries,
                             const uint32_t* __restrict__ node_tid, uint4* __restrict__ tnodes) {
  const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_entries) return;
  const uint64_t e = tl_node[k];
  const uint32_t slot = (uint32_t)e, kind = (uint32_t)(e >> 32);
  if (kind != K_TLAS_INTERNAL && kind != K_BLAS_INTERNAL) return;
  uint8_t* node = arena + (uint64_t)slot * 64u;
  const uint4* np = reinterpret_cast<const uint4*>(node);
  const uint4 a = np[0], b = np[1];
  const uint32_t own = node_tid[slot];
  uint32_t child = slot + a.w, m = 0, present = 0, leaf6 = 0, off = 0;
  uint64_t cbytes = 0;                                    // byte i: offset of child i | leaf << 6 | self-rooted << 7
  const uint64_t info6 = ((uint64_t)b.z << 16) | (b.y >> 16);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const uint32_t info = (uint32_t)(info6 >> (8 * i)) & 0xffu, sz = info & 3u;
    uint32_t self = 0;
    if (sz && child < n_slots) {
      const uint32_t t = node_tid[child];
      if (own != VSRT_NO_TID && t != VSRT_NO_TID && ((t ^ own) & VSRT_TID_MASK) == 0u) m |= 1u << i;
      self = (t & VSRT_TID_SELF_ROOTED) ? 0x80u : 0u;
    }
    const uint32_t leaf = (info & 0x3cu) ? 0x40u : 0u;
    node[22 + i] = (uint8_t)((info & 0x3fu) | leaf | self);
    cbytes |= (uint64_t)((off & 15u) | leaf | self) << (8 * i);
    if (sz) present |= 1u << i;
    if (leaf) leaf6 |= 1u << i;
    child += sz; off += sz;
  }
  node[17] = (uint8_t)m;
  if (tnodes) {
    const uint4 c = np[2], d = np[3];
    const uint32_t w[16] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w };
    auto byte_at = [&](int i) -> uint32_t { return (w[i >> 2] >> ((i & 3) * 8)) & 0xffu; };
    uint32_t t[16];
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = slot + a.w;
    t[4] = (uint32_t)cbytes; t[5] = (uint32_t)(cbytes >> 32) | (present << 16) | (leaf6 << 24);
    t[6] = ((byte_at(18) + 119u) & 0xffu) | (((byte_at(19) + 119u) & 0xffu) << 8) | (((byte_at(20) + 119u) & 0xffu) << 16) | (m << 24);
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const int lo = 28 + 12 * ax, hi = lo + 6;
      t[7 + 3 * ax] = byte_at(lo) | (byte_at(lo + 1) << 8) | (byte_at(lo + 2) << 16) | (byte_at(lo + 3) << 24);
      t[8 + 3 * ax] = byte_at(hi) | (byte_at(hi + 1) << 8) | (byte_at(hi + 2) << 16) | (byte_at(hi + 3) << 24);
      t[9 + 3 * ax] = byte_at(lo + 4) | (byte_at(lo + 5) << 8) | (byte_at(hi + 4) << 16) | (byte_at(hi + 5) << 24);
    }
    uint4* o = tnodes + (uint64_t)slot * 4u;
    o[0] = make_uint4(t[0], t[1], t[2], t[3]); o[1] = make_uint4(t[4], t[5], t[6], t[7]);
    o[2] = make_uint4(t[8], t[9], t[10], t[11]); o[3] = make_uint4(t[12], t[13], t[14], t[15]);
  }
}
// ---- remapBVHToTreeletLayout (:1473-1509) as a per-slot table: treelet t (ascending root order) starts at base + t * pitch,
// its root first, then its list entries in order; an entry keeps the mapping of the FIRST treelet (lowest index) that lists
// it but still advances the cursor of every later treelet (:1501-1503).  "First treelet that lists it" is an atomicMin over
// treelet indices, after which every treelet can lay out its own entries independently.
__global__ void k_remap_owner(const uint32_t* __restrict__ tl_root, const unsigned long long* __restrict__ tl_off, const uint64_t* __restrict__ tl_node,
                              uint32_t n_treelets, uint32_t* __restrict__ owner) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_treelets) return;
  atomicMin(owner + tl_root[t], t);
  for (unsigned long long k = tl_off[t]; k < tl_off[t + 1]; k++) atomicMin(owner + (uint32_t)tl_node[k], t);
}
__global__ void k_remap_assign(const uint32_t* __restrict__ tl_root, const unsigned long long* __restrict__ tl_off, const uint64_t* __restrict__ tl_node,
                               uint32_t n_treelets, const uint32_t* __restrict__ owner, unsigned long long base, unsigned long long pitch,
                               unsigned long long* __restrict__ remap) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_treelets) return;
  const uint32_t rslot = tl_root[t];
  const unsigned long long rnew = base + (unsigned long long)t * pitch;
  if (owner[rslot] == t) remap[rslot] = rnew;
  // root.first.size: 128 when the root is an instance leaf (:927), 64 otherwise
  unsigned long long cur = rnew + 64ull;
  for (unsigned long long k = tl_off[t]; k < tl_off[t + 1]; k++) if ((uint32_t)tl_node[k] == rslot && (uint32_t)(tl_node[k] >> 32) == K_INSTANCE) cur = rnew + 128ull;
  for (unsigned long long k = tl_off[t]; k < tl_off[t + 1]; k++) {
    const uint64_t e = tl_node[k]; const uint32_t slot = (uint32_t)e;
    if (slot == rslot) continue;
    if (owner[slot] == t) remap[slot] = cur;
    cur += ((uint32_t)(e >> 32) == K_INSTANCE) ? 128ull : 64ull;
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(errbuf, errcap, "%s: %s", #x, cudaGetErrorString(e_)); rc = VSRT_E_CUDA; goto done; } } while (0)

}  // namespace

int vsrt_launch_form_treelets(const ArenaView& av, uint32_t budget, cudaStream_t st, FormOutputs* out, FormResult* res,
                              uint32_t* err_flags_dev, char* errbuf, size_t errcap, uint8_t* tarena) {
  int rc = VSRT_OK;
  const uint32_t ns = av.n_slots, nw = (ns + 31) / 32;
  const uint32_t cap = budget / 64 + 2;
  uint32_t* claimed = nullptr; uint2* roots = nullptr; unsigned int* n_roots_d = nullptr; unsigned long long* scal = nullptr;
  uint32_t* r_count = nullptr; unsigned long long* r_off = nullptr; uint32_t* r_rank = nullptr;
  uint64_t* scratch = nullptr; uint64_t* store = nullptr; unsigned long long store_cap = 0; size_t scratch_roots = 0;
  uint32_t* popc = nullptr; unsigned long long* off64 = nullptr; void* scan_tmp = nullptr;
  uint32_t* prefix = nullptr; uint32_t* tl_root = nullptr; uint32_t* tl_count = nullptr; unsigned long long* tl_off = nullptr;
  uint64_t* tl_node = nullptr; uint32_t* node_tid = nullptr; uint4* tnodes = nullptr; unsigned long long* hot64 = nullptr; uint32_t* hot_keys = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint32_t n_roots = 1, begin = 0, h_err = 0;
  unsigned long long h_scal[3] = { 0, 0, 0 }, n_entries = 0;
  memset(out, 0, sizeof(*out)); memset(res, 0, sizeof(*res));
  if (budget < 192) { snprintf(errbuf, errcap, "max_treelet_size %u < 192 (an instance-leaf root is charged 128+64 bytes, reference asserts)", budget); return VSRT_E_BUDGET; }

  CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
  CK(cudaMalloc(&claimed, (size_t)nw * 4)); CK(cudaMemsetAsync(claimed, 0, (size_t)nw * 4, st));
  CK(cudaMalloc(&roots, (size_t)ns * sizeof(uint2)));
  CK(cudaMalloc(&n_roots_d, 4)); CK(cudaMalloc(&scal, 32)); CK(cudaMemsetAsync(scal, 0, 32, st)); CK(cudaMemsetAsync(scal + 2, 0xff, 4, st));   // [2] = {lowest = ~0, highest = 0} instance-leaf slot, [3] store cursor
  CK(cudaMalloc(&r_count, (size_t)ns * 4)); CK(cudaMalloc(&r_off, (size_t)ns * 8));
  // every node is listed once, plus the nodes that share a treelet with more than one instance of their BLAS
  store_cap = (unsigned long long)ns + ns / 4 + 256;
  CK(cudaMalloc(&store, (size_t)store_cap * 8));
  // scratch for the roots of one launch: at most 1 GiB, whatever the budget
  scratch_roots = std::max<size_t>(4096, ((size_t)1 << 30) / ((size_t)cap * 8));
  scratch_roots = std::min<size_t>(scratch_roots, (size_t)ns);
  CK(cudaMalloc(&scratch, scratch_roots * cap * 8));
  CK(cudaMemsetAsync(err_flags_dev, 0, 4, st));
  CK(cudaEventRecord(ev0, st));
  {
    // seed: the first treelet is keyed by the TLAS header (:847)
    const uint2 seed = make_uint2(av.tlas_slot, K_TLAS_HEADER); const unsigned int one = 1;
    const uint32_t bit = 1u << (av.tlas_slot & 31);
    CK(cudaMemcpyAsync(roots, &seed, sizeof(seed), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(n_roots_d, &one, 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(claimed + (av.tlas_slot >> 5), &bit, 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  {
    FormState fs = { claimed, roots, n_roots_d, scal, err_flags_dev, reinterpret_cast<uint32_t*>(scal + 2), store, scal + 3, store_cap, r_off };
    unsigned long long peak = (unsigned long long)scratch_roots * cap * 8;
    unsigned long long before[4] = { 0, 0, 0, 0 }, after[4] = { 0, 0, 0, 0 };     // [0] total_bvh, [3] store cursor: before / after a launch
    uint32_t n_roots_now = n_roots;
    while (begin < n_roots) {
      // one generation of roots, in launches of at most scratch_roots
      const uint32_t wave_end = n_roots;
      while (begin < wave_end) {
        const uint32_t count = (uint32_t)std::min<size_t>(wave_end - begin, scratch_roots);
        k_form_wave<<<(count + 127) / 128, 128, 0, st>>>(av, fs, begin, count, (int)budget, cap, scratch, r_count);
        CK(cudaGetLastError());
        // one read-back per launch: flags, roots discovered so far, totals and cursor
        CK(cudaMemcpyAsync(&h_err, err_flags_dev, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&n_roots_now, n_roots_d, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(after, scal, 32, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_err & EF_TRACE_CAP) {
          // the compact store is full: double it, rewind cursor and totals to where this launch started and repeat the launch
          uint64_t* bigger = nullptr; const unsigned long long ncap = store_cap * 2;
          CK(cudaMalloc(&bigger, (size_t)ncap * 8));
          CK(cudaMemcpyAsync(bigger, store, (size_t)before[3] * 8, cudaMemcpyDeviceToDevice, st));
          CK(cudaStreamSynchronize(st));
          cudaFree(store); store = bigger; store_cap = ncap; fs.store = store; fs.store_cap = store_cap;
          CK(cudaMemcpyAsync(scal + 3, &before[3], 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(scal, &before[0], 8, cudaMemcpyHostToDevice, st));
          const uint32_t cleared = h_err & ~(uint32_t)EF_TRACE_CAP;
          CK(cudaMemcpyAsync(err_flags_dev, &cleared, 4, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st));
          h_err = cleared;
          continue;
        }
        if (h_err & ~EF_NONFINITE) break;
        before[0] = after[0]; before[3] = after[3];
        begin += count;
      }
      if (h_err & ~EF_NONFINITE) break;
      n_roots = n_roots_now;
    }
    res->peak_scratch_bytes = peak + store_cap * 8;
  }
  res->nonfinite = (h_err & EF_NONFINITE) ? 1u : 0u;
  h_err &= ~(uint32_t)EF_NONFINITE;
  if (h_err) {
    rc = (h_err & EF_UNKNOWN_AS) ? VSRT_E_UNKNOWN_AS : (h_err & EF_BAD_BVH) ? VSRT_E_BAD_BVH : VSRT_E_BUDGET;
    snprintf(errbuf, errcap, "treelet formation rejected the arena (flags 0x%x): %s", h_err,
             rc == VSRT_E_UNKNOWN_AS ? "instance leaf references a BLAS that was never registered with vsrt_alloc_blas" :
             rc == VSRT_E_BAD_BVH ? "malformed BVH (child out of bounds, non-instance TLAS leaf or PrimitiveIndex1Delta != 0)" :
             "a node does not fit max_treelet_size (reference: assert(remaining_bytes >= 0))");
    goto done;
  }
  // ---- rank roots by address: exclusive popcount prefix over the bitmap
  CK(cudaMalloc(&popc, (size_t)nw * 4)); CK(cudaMalloc(&off64, ((size_t)std::max(nw, n_roots) + 1) * 8)); CK(cudaMalloc(&scan_tmp, vsrt_scan_tmp_bytes(std::max(nw, n_roots))));
  CK(cudaMalloc(&prefix, ((size_t)nw + 1) * 4));   // [nw] = lowest instance-leaf slot, for K1
  k_popc<<<(nw + 255) / 256, 256, 0, st>>>(claimed, nw, popc);
  if ((rc = vsrt_launch_scan(popc, nw, (uint64_t*)off64, scan_tmp, st)) != VSRT_OK) goto done;
  k_narrow<<<(nw + 255) / 256, 256, 0, st>>>(off64, nw, prefix);
  CK(cudaMalloc(&r_rank, (size_t)n_roots * 4)); CK(cudaMalloc(&tl_root, (size_t)n_roots * 4)); CK(cudaMalloc(&tl_count, (size_t)n_roots * 4));
  CK(cudaMalloc(&tl_off, ((size_t)n_roots + 1) * 8));
  CK(cudaMalloc(&hot64, (size_t)VSRT_HOT_N * 8)); CK(cudaMemsetAsync(hot64, 0xff, (size_t)VSRT_HOT_N * 8, st));
  CK(cudaMalloc(&hot_keys, (size_t)VSRT_HOT_N * 4));
  k_rank<<<(n_roots + 255) / 256, 256, 0, st>>>(roots, r_count, n_roots, claimed, prefix, r_rank, tl_root, tl_count, hot64);
  k_narrow<<<(VSRT_HOT_N + 255) / 256, 256, 0, st>>>(hot64, VSRT_HOT_N, hot_keys);   // low word = treelet index; an empty bucket keeps ~0 = VSRT_NO_TID
  if ((rc = vsrt_launch_scan(tl_count, n_roots, (uint64_t*)tl_off, scan_tmp, st)) != VSRT_OK) goto done;
  CK(cudaMemcpyAsync(&n_entries, tl_off + n_roots, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaMalloc(&tl_node, (size_t)std::max<unsigned long long>(n_entries, 1) * 8));
  CK(cudaMalloc(&node_tid, (size_t)ns * 4)); CK(cudaMemsetAsync(node_tid, 0, (size_t)ns * 4, st));
  k_gather<<<(n_roots + 127) / 128, 128, 0, st>>>(store, r_off, r_count, r_rank, n_roots, tl_off, tl_node, node_tid, scal + 1);
  k_fix_tid<<<(ns + 255) / 256, 256, 0, st>>>(node_tid, ns, claimed, prefix);
  // the arena copy is private to the context; its pad bytes are ours (data layout)
  // Two copies of the arena from here on: the Mesa-layout one (with K0's flag bits) that every other consumer reads, and K1's
  // traversal copy -- every slot verbatim, then the internal nodes overwritten in the traversal layout.  (Which of the two keeps
  // the allocation the arena was uploaded into makes no difference to K1: measured both ways.)
  tnodes = reinterpret_cast<uint4*>(tarena);     // allocated with the arena (vsrt_commit): no allocation inside the timed formation
  if (tnodes) CK(cudaMemcpyAsync(tnodes, av.base, (size_t)ns * 64, cudaMemcpyDeviceToDevice, st));
  if (n_entries) k_child_mask<<<(unsigned)((n_entries + 255) / 256), 256, 0, st>>>(const_cast<uint8_t*>(av.base), ns, tl_node, n_entries, node_tid, tnodes);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev1, st));
  CK(cudaMemcpyAsync(h_scal, scal, 24, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&res->ms, ev0, ev1));
  {
    // K1's stack entries name an instance leaf by its slot relative to the LOWEST instance leaf of the TLAS, in 23 bits: the
    // instance leaves of one TLAS (contiguous children of its internal nodes) must span less than 2^23 slots = 512 MiB
    const uint32_t lo = (uint32_t)h_scal[2], hi = (uint32_t)(h_scal[2] >> 32);
    res->inst_base = lo == 0xFFFFFFFFu ? 0u : lo;
    CK(cudaMemcpyAsync(prefix + nw, &res->inst_base, 4, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st));
    if (lo != 0xFFFFFFFFu && hi - lo >= 0x7FFFFFu) {
      snprintf(errbuf, errcap, "the instance leaves of this TLAS span %llu bytes; K1 addresses them in 23 bits of 64-byte slots (512 MiB, about 4 M instances)", (unsigned long long)(hi - lo) * 64ull);
      rc = VSRT_E_UNSUPPORTED; goto done;
    }
  }
  res->n_treelets = n_roots; res->n_entries = n_entries; res->n_mapped = h_scal[1]; res->total_bvh = h_scal[0];
  out->node_tid = node_tid; out->root_bits = claimed; out->root_prefix = prefix; out->tl_root = tl_root;
  out->tl_off = (uint64_t*)tl_off; out->tl_node = tl_node; out->tnodes = reinterpret_cast<uint8_t*>(tnodes);   // not owned: the context's traversal copy
  out->hot_keys = hot_keys; hot_keys = nullptr;
  node_tid = nullptr; claimed = nullptr; prefix = nullptr; tl_root = nullptr; tl_off = nullptr; tl_node = nullptr;
done:
  cudaFree(scratch); cudaFree(store);
  cudaFree(claimed); cudaFree(roots); cudaFree(n_roots_d); cudaFree(scal); cudaFree(r_count); cudaFree(r_off); cudaFree(r_rank);
  cudaFree(popc); cudaFree(off64); cudaFree(scan_tmp); cudaFree(prefix); cudaFree(tl_root); cudaFree(tl_count); cudaFree(tl_off);
  cudaFree(tl_node); cudaFree(node_tid); cudaFree(hot64); cudaFree(hot_keys);
  if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
  return rc;
}

int vsrt_launch_remap(const FormOutputs& fo, uint32_t n_treelets, uint32_t n_slots, uint64_t base, uint64_t pitch, uint64_t* remap_dev, cudaStream_t st) {
  uint32_t* owner = nullptr;
  if (cudaMalloc(&owner, (size_t)std::max(n_slots, 1u) * 4) != cudaSuccess) return VSRT_E_CUDA;
  cudaMemsetAsync(owner, 0xff, (size_t)n_slots * 4, st);
  cudaMemsetAsync(remap_dev, 0, (size_t)n_slots * 8, st);     // unmapped -> 0, like std::map::operator[] on a missing key
  if (n_treelets) {
    k_remap_owner<<<(n_treelets + 127) / 128, 128, 0, st>>>(fo.tl_root, (const unsigned long long*)fo.tl_off, fo.tl_node, n_treelets, owner);
    k_remap_assign<<<(n_treelets + 127) / 128, 128, 0, st>>>(fo.tl_root, (const unsigned long long*)fo.tl_off, fo.tl_node, n_treelets, owner, base, pitch,
                                                            (unsigned long long*)remap_dev);
  }
  const cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize(st);
  cudaFree(owner);
  return e == cudaSuccess ? VSRT_OK : VSRT_E_CUDA;
}
