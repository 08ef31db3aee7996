// gl64.cuh — Goldilocks field (p = 2^64 - 2^32 + 1) on the sm_100a integer pipes.
//
// Replaces plonky2_field 0.2.2 GoldilocksField (the type every reference call site instantiates:
// `type F = GoldilocksField`, city_rollup_core_worker/src/lib.rs:25-26).  Elements are raw u64,
// inputs may be non-canonical (>= p), stored outputs are canonical.
//
// Design: a 64x64 product is 4 32x32 IMAD(.WIDE/.HI) with the carries kept in predicate chains
// (mad.lo.cc/madc.hi.cc), and the 128->64 reduction uses 2^64 = 2^32-1, 2^96 = -1 (mod p) with both
// conditional corrections done branch-free through the carry flag (subc/addc masks).  ptxas emits
// 18 integer instructions per multiplication (7 on the FMA pipe, 11 on the ALU pipe).  (mul.lo + mul.hi pairs become
// IMAD + IMAD.HI; writing them as mul.wide.u32 — one IMAD.WIDE, 56 instructions less in the leaf-hash kernel — was
// measured in round 2 and changes nothing: leaf hashing 88.4 -> 89.1 ms at 2^20 x 135.)
// The sub.cc -> subc mask idiom is used only after subtract chains (where it is well defined).
#pragma once
#include <cstdint>

#define GL_P 0xFFFFFFFF00000001ull
#define GL_EPS 0xFFFFFFFFull

namespace gl {

__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// any u64 -> canonical.  a >= p  <=>  the high word is 2^32 - 1 and the low word is not 0, and then a - p = (lo - 1, 0):
// four instructions instead of the six of a 64-bit compare + subtract + select.
__device__ __forceinline__ uint64_t canon(uint64_t a) {
  uint32_t lo = (uint32_t)a, hi = (uint32_t)(a >> 32);
  const bool ge = (hi == 0xFFFFFFFFu) & (lo != 0u);
  return pack(lo - (ge ? 1u : 0u), ge ? 0u : hi);
}

// (a*b) mod p, any u64 inputs; result is a u64 congruent to the product, NOT necessarily < p.
__device__ __forceinline__ uint64_t mul_nc(uint64_t a, uint64_t b) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,m,tl,th;\n\t"
      // 128-bit product x3:x2:x1:x0
      "mul.lo.u32 x0, %2, %4;\n\t"
      "mul.hi.u32 x1, %2, %4;\n\t"
      "mul.lo.u32 x2, %3, %5;\n\t"
      "mul.hi.u32 x3, %3, %5;\n\t"
      "mad.lo.cc.u32 x1, %2, %5, x1;\n\t"
      "madc.hi.cc.u32 x2, %2, %5, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "mad.lo.cc.u32 x1, %3, %4, x1;\n\t"
      "madc.hi.cc.u32 x2, %3, %4, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      // t = (x1:x0) - x3; on borrow subtract EPS (2^64 = EPS mod p)
      "sub.cc.u32 tl, x0, x3;\n\t"
      "subc.cc.u32 th, x1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 tl, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      // r = t + x2*EPS; on carry add EPS = (c<<32) - c.  NOTE: the carry of an add chain must not be
      // turned into a mask with `subc` — ptxas keeps the raw adder carry in CC.CF, so add.cc -> subc
      // yields the complemented mask.  addc materialises c unambiguously and ptxas fuses the three
      // trailing ops into SEL + IADD3 + IADD3.X.
      "mad.lo.cc.u32 tl, x2, 0xFFFFFFFF, tl;\n\t"
      "madc.hi.cc.u32 th, x2, 0xFFFFFFFF, th;\n\t"
      "addc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      "add.u32 %1, th, m;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack(r0, r1);
}

// ---- variants for kernels that also keep the FP64 pipe busy (Poseidon) ------------------------------------
// On B200 the 32x32->64 multiplies (IMAD.WIDE / IMAD.HI, ~4.3 cycles per warp instruction) and DFMA (~2.3) do not
// overlap: tools/int32_peak.cu `mix_dfma_imadwide` runs at exactly the sum of the two, and in every ncu capture of the
// leaf-hash kernel fmaheavy% + fp64% = 98-99 %.  Where DFMA work runs beside the field multiplications, each wide
// multiply removed is worth two DFMAs, and the ALU pipe has room.  mul_nc_lw / sqr_nc are mul_nc with
//   * x2 * (2^32 - 1) formed as (x2 << 32) - x2 with carries (4 IADD3 instead of IADD3 + IMAD.HI + SEL), and
//   * for squares, ONE cross product a0*a1 doubled by a funnel shift instead of two accumulating IMAD.WIDE.
#define P2B_GL_REDUCE_LW                                                                                        \
  "sub.cc.u32 tl, x0, x3;\n\t"                                                                                  \
  "subc.cc.u32 th, x1, 0;\n\t"                                                                                  \
  "subc.u32 m, 0, 0;\n\t"                                                                                       \
  "sub.cc.u32 tl, tl, m;\n\t"                                                                                   \
  "subc.u32 th, th, 0;\n\t"                                                                                     \
  /* r = t + (x2 << 32) - x2: low word borrows b, high word gains u = x2 - b (>= 0: b = 1 implies x2 >= 1) */   \
  "sub.cc.u32 tl, tl, x2;\n\t"                                                                                  \
  "subc.u32 m, x2, 0;\n\t"                                                                                      \
  "add.cc.u32 th, th, m;\n\t"                                                                                   \
  "addc.u32 m, 0, 0;\n\t"                                                                                       \
  /* on carry add 2^32 - 1 */                                                                                   \
  "sub.cc.u32 %0, tl, m;\n\t"                                                                                   \
  "subc.u32 th, th, 0;\n\t"                                                                                     \
  "add.u32 %1, th, m;\n\t"
__device__ __forceinline__ uint64_t mul_nc_lw(uint64_t a, uint64_t b) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,m,tl,th;\n\t"
      "mul.lo.u32 x0, %2, %4;\n\t"
      "mul.hi.u32 x1, %2, %4;\n\t"
      "mul.lo.u32 x2, %3, %5;\n\t"
      "mul.hi.u32 x3, %3, %5;\n\t"
      "mad.lo.cc.u32 x1, %2, %5, x1;\n\t"
      "madc.hi.cc.u32 x2, %2, %5, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "mad.lo.cc.u32 x1, %3, %4, x1;\n\t"
      "madc.hi.cc.u32 x2, %3, %4, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t" P2B_GL_REDUCE_LW "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack(r0, r1);
}
// (a*a) mod p, any u64 input; result congruent, NOT necessarily < p
__device__ __forceinline__ uint64_t sqr_nc(uint64_t a) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,c0,c1,d0,d1,d2,m,tl,th;\n\t"
      "mul.lo.u32 x0, %2, %2;\n\t"
      "mul.hi.u32 x1, %2, %2;\n\t"
      "mul.lo.u32 x2, %3, %3;\n\t"
      "mul.hi.u32 x3, %3, %3;\n\t"
      "mul.lo.u32 c0, %2, %3;\n\t"
      "mul.hi.u32 c1, %2, %3;\n\t"
      "shl.b32 d0, c0, 1;\n\t"               // 2 * a0 * a1 = d2:d1:d0
      "shf.l.wrap.b32 d1, c0, c1, 1;\n\t"
      "shr.u32 d2, c1, 31;\n\t"
      "add.cc.u32 x1, x1, d0;\n\t"
      "addc.cc.u32 x2, x2, d1;\n\t"
      "addc.u32 x3, x3, d2;\n\t" P2B_GL_REDUCE_LW "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1));
  return pack(r0, r1);
}

__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) { return canon(mul_nc(a, b)); }

// (a*b + c) mod p, any u64 inputs (a*b + c < 2^128); result congruent, NOT necessarily < p.  One reduction for the
// product and the addend: the accumulate step of the linearised partial rounds (poseidon::coop_partial_rounds),
// where neither the running sum nor the product is canonical.
__device__ __forceinline__ uint64_t mad_nc(uint64_t a, uint64_t b, uint64_t c) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,m,tl,th;\n\t"
      "mul.lo.u32 x0, %2, %4;\n\t"
      "mul.hi.u32 x1, %2, %4;\n\t"
      "mul.lo.u32 x2, %3, %5;\n\t"
      "mul.hi.u32 x3, %3, %5;\n\t"
      "mad.lo.cc.u32 x1, %2, %5, x1;\n\t"
      "madc.hi.cc.u32 x2, %2, %5, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "mad.lo.cc.u32 x1, %3, %4, x1;\n\t"
      "madc.hi.cc.u32 x2, %3, %4, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "add.cc.u32 x0, x0, %6;\n\t"
      "addc.cc.u32 x1, x1, %7;\n\t"
      "addc.cc.u32 x2, x2, 0;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "sub.cc.u32 tl, x0, x3;\n\t"
      "subc.cc.u32 th, x1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 tl, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      "mad.lo.cc.u32 tl, x2, 0xFFFFFFFF, tl;\n\t"
      "madc.hi.cc.u32 th, x2, 0xFFFFFFFF, th;\n\t"
      "addc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      "add.u32 %1, th, m;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"((uint32_t)c), "r"((uint32_t)(c >> 32)));
  return pack(r0, r1);
}

// a + b mod p; requires a + b < 2^65 - 2^32 (true when at least one operand is canonical).
// Result is a u64 congruent to the sum, not necessarily canonical.
__device__ __forceinline__ uint64_t add_nc(uint64_t a, uint64_t b) {
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "add.cc.u32 %0, %2, %4;\n\t"
      "addc.cc.u32 %1, %3, %5;\n\t"
      "addc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "add.u32 %1, %1, m;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
  return pack(r0, r1);
}
// a - b mod p; requires b canonical (b - a <= p - 1).  Result congruent, not necessarily canonical.
__device__ __forceinline__ uint64_t sub_nc(uint64_t a, uint64_t b) {
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 m;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.cc.u32 %1, %3, %5;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
  return pack(r0, r1);
}
// Canonical in (BOTH operands), canonical out — the arithmetic of the constraint evaluators and the prover kernels,
// where every value is canonical (ncu on k_quotient at 2^16 rows: a quarter of all executed instructions were the
// generic canon() behind add / sub / mul).  a + b = a - (p - b): no intermediate exceeds 64 bits, and a borrow means
// "add p back", i.e. subtract 2^32 - 1 modulo 2^64.  7 instructions (was 12).
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 nl,nh,m;\n\t"
      "sub.cc.u32 nl, 1, %4;\n\t"  // p - b  (b < p: no borrow out)
      "subc.u32 nh, 0xFFFFFFFF, %5;\n\t"
      "sub.cc.u32 %0, %2, nl;\n\t"
      "subc.cc.u32 %1, %3, nh;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
  return pack(r0, r1);
}
// a - b for canonical a, b: sub_nc's result is then already canonical (no borrow: a - b < p; borrow: a - b + p < p)
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) { return sub_nc(a, b); }
__device__ __forceinline__ uint64_t neg(uint64_t a) { return a ? GL_P - a : 0; }  // a canonical

__device__ __forceinline__ uint64_t pow(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ uint64_t inv(uint64_t a) { return pow(a, GL_P - 2); }

// ---- quadratic extension F[X]/(X^2 - 7) (plonky2 QuadraticExtension<GoldilocksField>, W = 7) ----
struct ext2 {
  uint64_t c0, c1;
};
__device__ __forceinline__ ext2 ext_add(ext2 a, ext2 b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
__device__ __forceinline__ ext2 ext_sub(ext2 a, ext2 b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
// a0 * b0 + a1 * b1 mod p with ONE reduction (any u64 inputs, canonical result): the two 128-bit products are summed as a
// 129-bit integer x4:x3:x2:x1:x0 (2^128 = -2^32 mod p takes care of x4).  39 instructions instead of the 59 of
// add(mul, mul); exact integer arithmetic, so the field element is the same.
__device__ __forceinline__ uint64_t dot2(uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1) {
  uint32_t r0, r1, x4;
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,m,tl,th;\n\t"
      "mul.lo.u32 x0, %3, %5;\n\t"
      "mul.hi.u32 x1, %3, %5;\n\t"
      "mul.lo.u32 x2, %4, %6;\n\t"
      "mul.hi.u32 x3, %4, %6;\n\t"
      "mad.lo.cc.u32 x1, %3, %6, x1;\n\t"
      "madc.hi.cc.u32 x2, %3, %6, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      "mad.lo.cc.u32 x1, %4, %5, x1;\n\t"
      "madc.hi.cc.u32 x2, %4, %5, x2;\n\t"
      "addc.u32 x3, x3, 0;\n\t"
      // + a1 * b1 (the accumulate pattern of plonk::Acc::add_product)
      "mad.lo.cc.u32 x0, %7, %9, x0;\n\t"
      "madc.hi.cc.u32 x1, %7, %9, x1;\n\t"
      "madc.lo.cc.u32 x2, %8, %10, x2;\n\t"
      "madc.hi.cc.u32 x3, %8, %10, x3;\n\t"
      "addc.u32 %2, 0, 0;\n\t"
      "mad.lo.cc.u32 x1, %7, %10, x1;\n\t"
      "madc.hi.cc.u32 x2, %7, %10, x2;\n\t"
      "addc.cc.u32 x3, x3, 0;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      "mad.lo.cc.u32 x1, %8, %9, x1;\n\t"
      "madc.hi.cc.u32 x2, %8, %9, x2;\n\t"
      "addc.cc.u32 x3, x3, 0;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      // low 128 bits: exactly mul_nc's reduction
      "sub.cc.u32 tl, x0, x3;\n\t"
      "subc.cc.u32 th, x1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 tl, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      "mad.lo.cc.u32 tl, x2, 0xFFFFFFFF, tl;\n\t"
      "madc.hi.cc.u32 th, x2, 0xFFFFFFFF, th;\n\t"
      "addc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, tl, m;\n\t"
      "subc.u32 th, th, 0;\n\t"
      "add.u32 %1, th, m;\n\t"
      "}"
      : "=r"(r0), "=r"(r1), "=&r"(x4)  // x4 is written while inputs are still to be read: early clobber
      : "r"((uint32_t)a0), "r"((uint32_t)(a0 >> 32)), "r"((uint32_t)b0), "r"((uint32_t)(b0 >> 32)), "r"((uint32_t)a1),
        "r"((uint32_t)(a1 >> 32)), "r"((uint32_t)b1), "r"((uint32_t)(b1 >> 32)));
  // + x4 * 2^128 = - x4 * 2^32 (the sum of two products is below 2^129: x4 <= 1)
  return sub(canon(pack(r0, r1)), pack(0u, x4));
}
__device__ __forceinline__ ext2 ext_mul(ext2 a, ext2 b) {
  // c0 = a0 b0 + 7 a1 b1, c1 = a0 b1 + a1 b0: two dot products with one reduction each
  return {dot2(a.c0, b.c0, mul_nc(a.c1, b.c1), 7), dot2(a.c0, b.c1, a.c1, b.c0)};
}
__device__ __forceinline__ ext2 ext_scale(ext2 a, uint64_t s) { return {mul(a.c0, s), mul(a.c1, s)}; }
// sum of products a_i * b_i (any u64 operands) as an unreduced 160-bit integer (up to 2^32 terms), reduced once by
// value(): 2^128 = -2^32 (mod p).  13 instructions per term instead of the 33 of add(acc, mul(a, b)); exact.
struct Dot160 {
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
  __device__ __forceinline__ void add_product(uint64_t c, uint64_t a) {
    const uint32_t c0 = (uint32_t)c, c1 = (uint32_t)(c >> 32), a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %8, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "}"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(w4)
        : "r"(c0), "r"(c1), "r"(a0), "r"(a1));
  }
  __device__ __forceinline__ uint64_t value() const {  // canonical
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 m,tl,th;\n\t"
        "sub.cc.u32 tl, %2, %5;\n\t"
        "subc.cc.u32 th, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 tl, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "mad.lo.cc.u32 tl, %4, 0xFFFFFFFF, tl;\n\t"
        "madc.hi.cc.u32 th, %4, 0xFFFFFFFF, th;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "add.u32 %1, th, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    return sub(canon(pack(r0, r1)), pack(0u, w4));
  }
};


}  // namespace gl
