// vsrt_device.cuh -- device-side decode and ray math shared by the kernels.
//
// Bit-exactness contract: the reference is g++ -O3 on baseline x86-64 => scalar SSE2, no FMA, no
// reassociation, denormals kept.  The library is compiled with -fmad=false -prec-div=true -prec-sqrt=true
// -ftz=false; in addition every multiply that feeds an add is written with __fmul_rn/__fadd_rn so the operation
// order is explicit and contraction is impossible even if a flag is lost.  MIN/MAX are the reference's ternary
// macros (vulkan_ray_tracing.h:53-56): a NaN operand selects the SECOND argument, unlike fminf/fmaxf.
#pragma once
#include "vsrt_internal.h"

#define VS_DEV __device__ __forceinline__

VS_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
VS_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
VS_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
VS_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
VS_DEV float rmin(float a, float b) { return (a < b) ? a : b; }
VS_DEV float rmax(float a, float b) { return (a > b) ? a : b; }

struct Node64 { uint32_t w[16]; };

VS_DEV Node64 load_node(const uint8_t* base, uint32_t slot) {
  const uint4* p = reinterpret_cast<const uint4*>(base + (uint64_t)slot * 64u);
  Node64 n;
  uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  n.w[0] = a.x; n.w[1] = a.y; n.w[2] = a.z; n.w[3] = a.w; n.w[4] = b.x; n.w[5] = b.y; n.w[6] = b.z; n.w[7] = b.w;
  n.w[8] = c.x; n.w[9] = c.y; n.w[10] = c.z; n.w[11] = c.w; n.w[12] = d.x; n.w[13] = d.y; n.w[14] = d.z; n.w[15] = d.w;
  return n;
}
// Same, but the four loads are issued here and now: the compiler otherwise sinks three of them below the first branch on
// the node's contents (leaf type), which costs a BLAS-leaf visit two dependent memory round trips instead of one.
VS_DEV Node64 load_node_now(const uint8_t* base, uint32_t slot) {
  const uint4* p = reinterpret_cast<const uint4*>(base + (uint64_t)slot * 64u);
  Node64 n;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n.w[0]), "=r"(n.w[1]), "=r"(n.w[2]), "=r"(n.w[3]) : "l"(p));
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n.w[4]), "=r"(n.w[5]), "=r"(n.w[6]), "=r"(n.w[7]) : "l"(p + 1));
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n.w[8]), "=r"(n.w[9]), "=r"(n.w[10]), "=r"(n.w[11]) : "l"(p + 2));
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n.w[12]), "=r"(n.w[13]), "=r"(n.w[14]), "=r"(n.w[15]) : "l"(p + 3));
  return n;
}
// byte i (0..63) of a node held in registers; i must be a compile-time constant after unrolling
VS_DEV uint32_t node_byte(const Node64& n, int i) { return (n.w[i >> 2] >> ((i & 3) * 8)) & 0xffu; }

// GEN_RT_BVH_INTERNAL_NODE fields (reference util.h:134-207; wire offsets SURVEY A.1)
VS_DEV int32_t node_child_offset(const Node64& n) { return (int32_t)n.w[3]; }
VS_DEV uint32_t node_child_info(const Node64& n, int i) { return node_byte(n, 22 + i) & 0x3fu; }   // size = &3, type = >>2
// 2^(e-8) as an exact float for the int8 exponent byte at `idx` (e-8 in [-136,119]); (float)q * 2^k is exact and
// therefore equal to the reference's ldexpf((float)q, k)  (util.h:499-508).
VS_DEV float node_scale(const Node64& n, int idx) {
  const int k = (int)(int8_t)node_byte(n, idx) - 8;
  // normal floats for k >= -126 (every sane BVH); denormal powers of two below that
  const uint32_t bits = (k >= -126) ? ((uint32_t)(k + 127) << 23) : (1u << ((k + 149) & 31));
  return __uint_as_float(bits);
}

struct Ray8 { float ox, oy, oz, dx, dy, dz, tmin, tmax; };
struct Idir { float x, y, z; };

// calculate_idir, vulkan_ray_tracing.cc:220-235
VS_DEV float idir1(float d) {
  const float ooeps = 8.271806125530277e-25f;   // 2^-80
  return fdiv(1.0f, (fabsf(d) > ooeps) ? d : copysignf(ooeps, d));
}
VS_DEV Idir calc_idir(const Ray8& r) { Idir i; i.x = idir1(r.dx); i.y = idir1(r.dy); i.z = idir1(r.dz); return i; }

// ray_box_test + get_t_bound + magic_max7/magic_min7, vulkan_ray_tracing.cc:183-257
VS_DEV bool ray_box(float lox, float loy, float loz, float hix, float hiy, float hiz, const Idir& id, const Ray8& r, float& thit) {
  float lx = fmul(fsub(lox, r.ox), id.x), ly = fmul(fsub(loy, r.oy), id.y), lz = fmul(fsub(loz, r.oz), id.z);
  float hx = fmul(fsub(hix, r.ox), id.x), hy = fmul(fsub(hiy, r.oy), id.y), hz = fmul(fsub(hiz, r.oz), id.z);
  float t1 = rmax(rmin(lx, hx), r.tmin), t2 = rmax(rmin(ly, hy), t1), mn = rmax(rmin(lz, hz), t2);
  float u1 = rmin(rmax(lx, hx), r.tmax), u2 = rmin(rmax(ly, hy), u1), mx = rmin(rmax(lz, hz), u2);
  thit = mn;
  return mn <= mx;
}

// Exact (float)byte without the quarter-rate I2F pipe: PRMT builds 0x4B0000qq == 2^23 + q, the subtraction is exact.
VS_DEV float byte_to_float(const Node64& n, int i) {
  return fsub(__uint_as_float(__byte_perm(n.w[i >> 2], 0x4B000000u, 0x7540u | (uint32_t)(i & 3))), 8388608.0f);
}
// lo = Origin + ldexpf((float)q, exp-8) (util.h:499-508).  q * 2^k is exact, so the single rounding of the
// reference's add is exactly the single rounding of one FMA.
VS_DEV float dequant(const Node64& n, int i, float scale, float org) { return __fmaf_rn(byte_to_float(n, i), scale, org); }

// ray_box_test with IEEE min/max (one FMNMX each) instead of the ternary macros.  Identical results whenever no
// operand is NaN: the only other difference is the sign of a zero result, which no later comparison can observe.
// Callers use it only for rays and nodes whose coordinates were checked to be NaN-free (see `exact` below).
VS_DEV bool ray_box_fast(float lox, float loy, float loz, float hix, float hiy, float hiz, const Idir& id, const Ray8& r, float& thit) {
  float lx = fmul(fsub(lox, r.ox), id.x), ly = fmul(fsub(loy, r.oy), id.y), lz = fmul(fsub(loz, r.oz), id.z);
  float hx = fmul(fsub(hix, r.ox), id.x), hy = fmul(fsub(hiy, r.oy), id.y), hz = fmul(fsub(hiz, r.oz), id.z);
  float mn = fmaxf(fminf(lz, hz), fmaxf(fminf(ly, hy), fmaxf(fminf(lx, hx), r.tmin)));
  float mx = fminf(fmaxf(lz, hz), fminf(fmaxf(ly, hy), fminf(fmaxf(lx, hx), r.tmax)));
  thit = mn;
  return mn <= mx;
}

// ---- packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per issued instruction,
// each half bit-identical to the scalar __fadd_rn / __fmul_rn / __fmaf_rn (no FTZ, no contraction).
VS_DEV uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
VS_DEV void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
VS_DEV uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
VS_DEV uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
VS_DEV uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// Tests the six child boxes of an internal node; returns the hit mask after the reference's cull
// `thit >= min_thit * tMult` (:1791,:1989,:2537,:2725).  `cull` = min_thit * tMult computed by the caller.
// EXACT = true keeps the reference's ternary MIN/MAX (NaN picks the second operand); it is taken for rays or
// arenas with non-finite coordinates, where a NaN can reach the slab test, and for arenas in which a present child has a
// quantised lower bound above its upper bound (the fast path reads near/far planes off the ray's direction signs).  Straight-line code: all six slots are
// evaluated and empty slots (ChildSize == 0) are masked out at the end.
template <bool EXACT>
VS_DEV uint32_t test_children(const Node64& n, const Ray8& r, const Idir& id, float cull, uint32_t magic16) {
  const float ox = __uint_as_float(n.w[0]), oy = __uint_as_float(n.w[1]), oz = __uint_as_float(n.w[2]);
  // 2^(e-8) per axis.  Exponents below -126 (denormal scales, no sane BVH) need the general form: K0 sends an arena that has
  // one down the EXACT path, so the hot path builds the float from the exponent byte alone: (e + 119) << 23, e >= -118
  float sx, sy, sz;
  if (EXACT) { sx = node_scale(n, 18); sy = node_scale(n, 19); sz = node_scale(n, 20); }
  else {
    sx = __uint_as_float(((node_byte(n, 18) + 119u) & 0xffu) << 23); sy = __uint_as_float(((node_byte(n, 19) + 119u) & 0xffu) << 23);
    sz = __uint_as_float(((node_byte(n, 20) + 119u) & 0xffu) << 23);
  }
  uint32_t mask = 0;
  if (EXACT) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const float lox = dequant(n, 28 + i, sx, ox), hix = dequant(n, 34 + i, sx, ox);
      const float loy = dequant(n, 40 + i, sy, oy), hiy = dequant(n, 46 + i, sy, oy);
      const float loz = dequant(n, 52 + i, sz, oz), hiz = dequant(n, 58 + i, sz, oz);
      float th;
      const bool h = ray_box(lox, loy, loz, hix, hiy, hiz, id, r, th);
      mask |= (h && !(th >= cull) && (node_byte(n, 22 + i) & 3u) != 0u) ? (1u << i) : 0u;
    }
  } else {
    // Same operations in the same order as ray_box_fast(dequant(..)) for every child, issued two children at a time:
    //  * NEAR / FAR instead of min / max per axis.  t_lo = ((q_lo * s + o) - ro) * idir and t_hi are monotone in q (every
    //    step rounds monotonically, s > 0), so with q_lo <= q_hi -- K0 checks it for every present child and sends arenas that
    //    break it down the EXACT path -- min(t_lo, t_hi) IS t_lo when idir >= 0 and t_hi when idir < 0 (no NaN can arise: idir
    //    is finite and non-zero by calculate_idir, node and ray coordinates were checked finite).  The six min/max per child
    //    become a choice of BYTES made once per node and axis: the axis' twelve bound bytes (lower 0..5, upper 0..5) are
    //    regrouped by four PRMTs with per-ray selectors into near(0..3), far(0..3) and near(4,5)|far(4,5);
    //  * one PRMT per byte builds 2^23 + q as fp32 and a packed add of -2^23 gives (float)q exactly; de-quantisation, origin
    //    subtraction and the multiply by idir run as packed f32x2 (a - b is written a + (-b), the same IEEE operation);
    //  * hit = (mn <= mx) && !(mn >= cull) is read from sign bits: mx - mn is negative exactly when mn > mx, mn - cull is
    //    negative exactly when mn < cull (a zero or NaN difference has a clear sign, like the comparisons being false; the
    //    only NaN possible here is inf - inf, which FADD returns as the positive default NaN).  Children are visited
    //    5..0 so that a funnel shift leaves child i's bit at position i.
    const uint64_t s_x = pk2(sx, sx), s_y = pk2(sy, sy), s_z = pk2(sz, sz), o_x = pk2(ox, ox), o_y = pk2(oy, oy), o_z = pk2(oz, oz);
    const uint64_t nr_x = pk2(-r.ox, -r.ox), nr_y = pk2(-r.oy, -r.oy), nr_z = pk2(-r.oz, -r.oz);
    const uint64_t id_x = pk2(id.x, id.x), id_y = pk2(id.y, id.y), id_z = pk2(id.z, id.z);
    const uint32_t magic = magic16 ^ 0x2F646464u;   // 0x4B000000 in a register, so the byte selector can be the PRMT's immediate
    const uint64_t m23 = pk2(-8388608.0f, -8388608.0f);
    // near4 = bytes of children 0..3 on the side the ray enters, far4 on the side it leaves, nf2 = near(4,5) | far(4,5)
    uint32_t near4[3], far4[3], nf2[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const bool neg = (__float_as_uint(ax == 0 ? id.x : (ax == 1 ? id.y : id.z)) >> 31) != 0u;
      const uint32_t wA = n.w[7 + 3 * ax], wB = n.w[8 + 3 * ax], wC = n.w[9 + 3 * ax];   // lower 0..3 | lower 4,5 upper 0,1 | upper 2..5
      const uint32_t up4 = __byte_perm(wB, wC, 0x5432);                                   // upper 0..3
      const uint32_t sn = neg ? 0x7654u : 0x3210u;
      near4[ax] = __byte_perm(wA, up4, sn); far4[ax] = __byte_perm(wA, up4, sn ^ 0x4444u);
      nf2[ax] = __byte_perm(wB, wC, neg ? 0x1076u : 0x7610u);
    }
#define VS_B2(w_, k_) pk2(__uint_as_float(__byte_perm((w_), magic, 0x7540u + (k_))), __uint_as_float(__byte_perm((w_), magic, 0x7541u + (k_))))
#define VS_T2(w_, k_, s_, o_, nr_, id_) mul2(add2(fma2(add2(VS_B2(w_, k_), m23), s_, o_), nr_), id_)
#pragma unroll
    for (int pp = 2; pp >= 0; pp--) {
      float nx[2], fx[2], ny[2], fy[2], nz[2], fz[2];
      if (pp == 2) {
        upk2(VS_T2(nf2[0], 0, s_x, o_x, nr_x, id_x), nx[0], nx[1]); upk2(VS_T2(nf2[0], 2, s_x, o_x, nr_x, id_x), fx[0], fx[1]);
        upk2(VS_T2(nf2[1], 0, s_y, o_y, nr_y, id_y), ny[0], ny[1]); upk2(VS_T2(nf2[1], 2, s_y, o_y, nr_y, id_y), fy[0], fy[1]);
        upk2(VS_T2(nf2[2], 0, s_z, o_z, nr_z, id_z), nz[0], nz[1]); upk2(VS_T2(nf2[2], 2, s_z, o_z, nr_z, id_z), fz[0], fz[1]);
      } else {
        upk2(VS_T2(near4[0], 2 * pp, s_x, o_x, nr_x, id_x), nx[0], nx[1]); upk2(VS_T2(far4[0], 2 * pp, s_x, o_x, nr_x, id_x), fx[0], fx[1]);
        upk2(VS_T2(near4[1], 2 * pp, s_y, o_y, nr_y, id_y), ny[0], ny[1]); upk2(VS_T2(far4[1], 2 * pp, s_y, o_y, nr_y, id_y), fy[0], fy[1]);
        upk2(VS_T2(near4[2], 2 * pp, s_z, o_z, nr_z, id_z), nz[0], nz[1]); upk2(VS_T2(far4[2], 2 * pp, s_z, o_z, nr_z, id_z), fz[0], fz[1]);
      }
#pragma unroll
      for (int k = 1; k >= 0; k--) {
        const float mn = fmaxf(fmaxf(nx[k], ny[k]), fmaxf(nz[k], r.tmin));
        const float mx = fminf(fminf(fx[k], fy[k]), fminf(fz[k], r.tmax));
        const uint32_t sb = __float_as_uint(fsub(mn, cull)) & ~__float_as_uint(fsub(mx, mn));   // bit 31 = hit
        mask = __funnelshift_l(sb, mask, 1);
      }
    }
#undef VS_T2
#undef VS_B2
    // empty slots (ChildSize == 0) never hit: bit 0 of each info byte = size != 0, gathered into six bits by a multiply
    const uint32_t lo4 = __byte_perm(n.w[5], n.w[6], 0x5432), hi2 = n.w[6] >> 16;                  // bytes 22..25 | 26,27
    const uint32_t nz4 = (lo4 | (lo4 >> 1)) & 0x01010101u, nz2 = (hi2 | (hi2 >> 1)) & 0x0101u;
    const uint32_t present = (((nz4 * 0x00204081u) >> 21) & 15u) | (((nz2 * 0x00204081u) >> 17) & 0x30u);
    mask &= present;
  }
  return mask;
}
// The hot slab test over a node in K1's TRAVERSAL LAYOUT (treelets.cu, k_child_mask): the same operations on the same values as
// test_children<false> -- only the byte shuffles that prepare them are gone.  Near / far planes of children 0..3 are whole words
// picked by the sign of the ray direction, those of children 4,5 one word rotated by 16 bits; the scale's exponent bits are
// stored ready to shift; the present mask is a field.
VS_DEV uint32_t test_children_t(const Node64& n, const Ray8& r, const Idir& id, float cull, uint32_t magic16) {
  const float ox = __uint_as_float(n.w[0]), oy = __uint_as_float(n.w[1]), oz = __uint_as_float(n.w[2]);
  const float sx = __uint_as_float((n.w[6] & 0xffu) << 23), sy = __uint_as_float((n.w[6] << 15) & 0x7F800000u), sz = __uint_as_float((n.w[6] << 7) & 0x7F800000u);
  const uint64_t s_x = pk2(sx, sx), s_y = pk2(sy, sy), s_z = pk2(sz, sz), o_x = pk2(ox, ox), o_y = pk2(oy, oy), o_z = pk2(oz, oz);
  const uint64_t nr_x = pk2(-r.ox, -r.ox), nr_y = pk2(-r.oy, -r.oy), nr_z = pk2(-r.oz, -r.oz);
  const uint64_t id_x = pk2(id.x, id.x), id_y = pk2(id.y, id.y), id_z = pk2(id.z, id.z);
  const uint32_t magic = magic16 ^ 0x2F646464u;   // 0x4B000000 in a register, so the byte selector can be the PRMT's immediate
  const uint64_t m23 = pk2(-8388608.0f, -8388608.0f);
  uint32_t near4[3], far4[3], nf2[3];
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const bool neg = (int32_t)__float_as_uint(ax == 0 ? id.x : (ax == 1 ? id.y : id.z)) < 0;
    const uint32_t lo4 = n.w[7 + 3 * ax], up4 = n.w[8 + 3 * ax], w45 = n.w[9 + 3 * ax];   // lower 0..3 | upper 0..3 | lower 4,5 upper 4,5
    near4[ax] = neg ? up4 : lo4; far4[ax] = neg ? lo4 : up4;
    nf2[ax] = __funnelshift_l(w45, w45, neg ? 16u : 0u);                                    // near(4,5) | far(4,5)
  }
#define VS_B2(w_, k_) pk2(__uint_as_float(__byte_perm((w_), magic, 0x7540u + (k_))), __uint_as_float(__byte_perm((w_), magic, 0x7541u + (k_))))
#define VS_T2(w_, k_, s_, o_, nr_, id_) mul2(add2(fma2(add2(VS_B2(w_, k_), m23), s_, o_), nr_), id_)
  uint32_t mask = 0;
#pragma unroll
  for (int pp = 2; pp >= 0; pp--) {
    float nx[2], fx[2], ny[2], fy[2], nz[2], fz[2];
    if (pp == 2) {
      upk2(VS_T2(nf2[0], 0, s_x, o_x, nr_x, id_x), nx[0], nx[1]); upk2(VS_T2(nf2[0], 2, s_x, o_x, nr_x, id_x), fx[0], fx[1]);
      upk2(VS_T2(nf2[1], 0, s_y, o_y, nr_y, id_y), ny[0], ny[1]); upk2(VS_T2(nf2[1], 2, s_y, o_y, nr_y, id_y), fy[0], fy[1]);
      upk2(VS_T2(nf2[2], 0, s_z, o_z, nr_z, id_z), nz[0], nz[1]); upk2(VS_T2(nf2[2], 2, s_z, o_z, nr_z, id_z), fz[0], fz[1]);
    } else {
      upk2(VS_T2(near4[0], 2 * pp, s_x, o_x, nr_x, id_x), nx[0], nx[1]); upk2(VS_T2(far4[0], 2 * pp, s_x, o_x, nr_x, id_x), fx[0], fx[1]);
      upk2(VS_T2(near4[1], 2 * pp, s_y, o_y, nr_y, id_y), ny[0], ny[1]); upk2(VS_T2(far4[1], 2 * pp, s_y, o_y, nr_y, id_y), fy[0], fy[1]);
      upk2(VS_T2(near4[2], 2 * pp, s_z, o_z, nr_z, id_z), nz[0], nz[1]); upk2(VS_T2(far4[2], 2 * pp, s_z, o_z, nr_z, id_z), fz[0], fz[1]);
    }
#pragma unroll
    for (int k = 1; k >= 0; k--) {
      const float mn = fmaxf(fmaxf(nx[k], ny[k]), fmaxf(nz[k], r.tmin));
      const float mx = fminf(fminf(fx[k], fy[k]), fminf(fz[k], r.tmax));
      const uint32_t sb = __float_as_uint(fsub(mn, cull)) & ~__float_as_uint(fsub(mx, mn));   // bit 31 = hit
      mask = __funnelshift_l(sb, mask, 1);
    }
  }
#undef VS_T2
#undef VS_B2
  return mask & (n.w[5] >> 16);      // present children only (the leaf mask above bit 7 cannot reach the six mask bits: mask < 64)
}
VS_DEV bool finite3(float a, float b, float c) { return (fabsf(a) <= 3.402823466e38f) && (fabsf(b) <= 3.402823466e38f) && (fabsf(c) <= 3.402823466e38f); }
// true if a NaN could reach the slab test for this ray: any non-finite origin/direction, NaN tmin/tmax
VS_DEV bool ray_needs_exact(const Ray8& r) { return !(finite3(r.ox, r.oy, r.oz) && finite3(r.dx, r.dy, r.dz) && r.tmin == r.tmin && r.tmax == r.tmax); }

// Instance leaf -> object-space ray.  make_transformed_ray (:168-181) with float4x4::operator* (util.h:47-55):
// res[i] = 0 + m[0][i]*v0 + m[1][i]*v1 + m[2][i]*v2 + m[3][i]*v3, W2O = wire A[0..8] rows + B[9..11] as row 3
// (SURVEY A.1 matrix trap), column 3 = (0,0,0,1).
struct InstCtx { Ray8 ray; Idir idir; float tmult; uint32_t inst_slot; bool exact; };
VS_DEV void make_object_ray(const uint8_t* base, uint32_t leaf_slot, const Ray8& w, InstCtx& c) {
  const float* A = reinterpret_cast<const float*>(base + (uint64_t)leaf_slot * 64u + 16u);
  const float* B = reinterpret_cast<const float*>(base + (uint64_t)leaf_slot * 64u + 80u);
  float m[4][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) m[r][k] = __ldg(A + 3 * r + k);
#pragma unroll
  for (int k = 0; k < 3; k++) m[3][k] = __ldg(B + 9 + k);
  float o[4], d[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    o[i] = fadd(fadd(fadd(fadd(0.0f, fmul(m[0][i], w.ox)), fmul(m[1][i], w.oy)), fmul(m[2][i], w.oz)), fmul(m[3][i], 1.0f));
    d[i] = fadd(fadd(fadd(fadd(0.0f, fmul(m[0][i], w.dx)), fmul(m[1][i], w.dy)), fmul(m[2][i], w.dz)), fmul(m[3][i], 0.0f));
  }
  o[3] = fadd(fadd(fadd(fadd(0.0f, fmul(0.0f, w.ox)), fmul(0.0f, w.oy)), fmul(0.0f, w.oz)), fmul(1.0f, 1.0f));
  // the divide by w of make_transformed_ray (:172): w is exactly 1.0f for the affine matrices an instance leaf can hold (column 3 =
  // (0,0,0,1): 0*x + 0*y + 0*z + 1*1), and x / 1.0f == x in IEEE arithmetic -- the three divisions run only if it ever is not
  c.ray.ox = o[0]; c.ray.oy = o[1]; c.ray.oz = o[2];
  if (o[3] != 1.0f) { c.ray.ox = fdiv(o[0], o[3]); c.ray.oy = fdiv(o[1], o[3]); c.ray.oz = fdiv(o[2], o[3]); }
  float norm = __fsqrt_rn(fadd(fadd(fmul(d[0], d[0]), fmul(d[1], d[1])), fmul(d[2], d[2])));
  c.tmult = norm;
  c.ray.dx = fdiv(d[0], norm); c.ray.dy = fdiv(d[1], norm); c.ray.dz = fdiv(d[2], norm);
  c.ray.tmin = fmul(w.tmin, norm); c.ray.tmax = fmul(w.tmax, norm);
  c.idir = calc_idir(c.ray);
  c.inst_slot = leaf_slot;
  c.exact = ray_needs_exact(c.ray);
}

VS_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz)); }

// mt_ray_triangle_test, vulkan_ray_tracing.cc:3089-3111 (no epsilon, no culling); leaf = quad leaf in registers
VS_DEV bool ray_tri(const Node64& q, const Ray8& r, float& thit) {
  const float p0x = __uint_as_float(q.w[4]), p0y = __uint_as_float(q.w[5]), p0z = __uint_as_float(q.w[6]);
  const float e1x = fsub(__uint_as_float(q.w[7]), p0x), e1y = fsub(__uint_as_float(q.w[8]), p0y), e1z = fsub(__uint_as_float(q.w[9]), p0z);
  const float e2x = fsub(__uint_as_float(q.w[10]), p0x), e2y = fsub(__uint_as_float(q.w[11]), p0y), e2z = fsub(__uint_as_float(q.w[12]), p0z);
  // pvec = cross(dir, v0v2)   (vector-math.cc:41-44)
  const float px = fsub(fmul(r.dy, e2z), fmul(r.dz, e2y)), py = fsub(fmul(r.dz, e2x), fmul(r.dx, e2z)), pz = fsub(fmul(r.dx, e2y), fmul(r.dy, e2x));
  const float det = dot3(e1x, e1y, e1z, px, py, pz);
  const float idet = fdiv(1.0f, det);
  const float tx = fsub(r.ox, p0x), ty = fsub(r.oy, p0y), tz = fsub(r.oz, p0z);
  const float u = fmul(dot3(tx, ty, tz, px, py, pz), idet);
  if (u < 0 || u > 1) return false;
  // qvec = cross(tvec, v0v1)
  const float qx = fsub(fmul(ty, e1z), fmul(tz, e1y)), qy = fsub(fmul(tz, e1x), fmul(tx, e1z)), qz = fsub(fmul(tx, e1y), fmul(ty, e1x));
  const float v = fmul(dot3(r.dx, r.dy, r.dz, qx, qy, qz), idet);
  if (v < 0 || fadd(u, v) > 1) return false;
  thit = fmul(dot3(e2x, e2y, e2z, qx, qy, qz), idet);
  return true;
}

// Barycentric, vulkan_ray_tracing.cc:3113-3130 -> {v, w, u}
VS_DEV void barycentric(const Node64& q, float px, float py, float pz, float out[3]) {
  const float ax = __uint_as_float(q.w[4]), ay = __uint_as_float(q.w[5]), az = __uint_as_float(q.w[6]);
  const float v0x = fsub(__uint_as_float(q.w[7]), ax), v0y = fsub(__uint_as_float(q.w[8]), ay), v0z = fsub(__uint_as_float(q.w[9]), az);
  const float v1x = fsub(__uint_as_float(q.w[10]), ax), v1y = fsub(__uint_as_float(q.w[11]), ay), v1z = fsub(__uint_as_float(q.w[12]), az);
  const float v2x = fsub(px, ax), v2y = fsub(py, ay), v2z = fsub(pz, az);
  const float d00 = dot3(v0x, v0y, v0z, v0x, v0y, v0z), d01 = dot3(v0x, v0y, v0z, v1x, v1y, v1z), d11 = dot3(v1x, v1y, v1z, v1x, v1y, v1z);
  const float d20 = dot3(v2x, v2y, v2z, v0x, v0y, v0z), d21 = dot3(v2x, v2y, v2z, v1x, v1y, v1z);
  const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
  const float v = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
  const float w = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
  out[0] = v; out[1] = w; out[2] = fsub(fsub(1.0f, v), w);
}

// ---- address translation between host addresses and packed-arena slots ----
VS_DEV uint32_t span_of_slot(const ArenaView& av, uint32_t slot) {
  uint32_t lo = 0, hi = av.n_spans;
  while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (av.spans[m].slot0 <= slot) lo = m; else hi = m; }
  return lo;
}
VS_DEV uint64_t slot_to_host(const ArenaView& av, uint32_t slot) {
  const Span& s = av.spans[av.n_spans == 1 ? 0 : span_of_slot(av, slot)];
  return s.host + (uint64_t)(slot - s.slot0) * 64u;
}
// returns false if the host address is outside every registered span or not 64-byte aligned
VS_DEV bool host_to_slot(const ArenaView& av, uint64_t host, uint32_t& slot) {
  uint32_t lo = 0, hi = av.n_spans;
  while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (av.spans[m].host <= host) lo = m; else hi = m; }
  const Span& s = av.spans[lo];
  if (host < s.host || host - s.host >= s.size || ((host - s.host) & 63u)) return false;
  slot = s.slot0 + (uint32_t)((host - s.host) >> 6);
  return true;
}
// blas_addr_map lookup by header slot; false if the header was never registered with allocBLAS
VS_DEV bool blas_delta_of(const ArenaView& av, uint32_t hdr_slot, int64_t& delta) {
  uint32_t lo = 0, hi = av.n_blas;
  while (lo < hi) { uint32_t m = (lo + hi) >> 1; uint32_t s = av.blas[m].hdr_slot; if (s == hdr_slot) { delta = av.blas[m].delta; return true; } if (s < hdr_slot) lo = m + 1; else hi = m; }
  return false;
}
// instance leaf -> slot of the BLAS header it references (BVHAddress is relative to the leaf, :1902)
VS_DEV bool instance_blas_header(const ArenaView& av, uint32_t leaf_slot, uint32_t& hdr_slot) {
  const uint64_t rel = __ldg(reinterpret_cast<const unsigned long long*>(av.base + (uint64_t)leaf_slot * 64u + 64u));
  if (rel == 0) return false;   // reference: assert(instanceLeaf.BVHAddress != NULL)
  if (av.n_spans == 1 && (rel & 63u) == 0) {
    int64_t s = (int64_t)leaf_slot + ((int64_t)rel >> 6);
    if (s < 0 || s >= (int64_t)av.n_slots) return false;
    hdr_slot = (uint32_t)s; return true;
  }
  return host_to_slot(av, slot_to_host(av, leaf_slot) + rel, hdr_slot);
}
// BLAS header -> slot of its root internal node (RootNodeOffset, bytes from the header)
VS_DEV bool header_root(const ArenaView& av, uint32_t hdr_slot, uint32_t& root_slot) {
  const uint64_t off = __ldg(reinterpret_cast<const unsigned long long*>(av.base + (uint64_t)hdr_slot * 64u));
  if (off & 63u) return false;
  uint64_t s = (uint64_t)hdr_slot + (off >> 6);
  if (s >= av.n_slots) return false;
  root_slot = (uint32_t)s; return true;
}
VS_DEV uint32_t root_rank(const TreeletView& tv, uint32_t slot) {   // treelet index of the treelet ROOTED at slot, or NO_TID
  uint32_t w = __ldg(tv.root_bits + (slot >> 5)), b = 1u << (slot & 31);
  if (!(w & b)) return VSRT_NO_TID;
  return __ldg(tv.root_prefix + (slot >> 5)) + __popc(w & (b - 1));
}
