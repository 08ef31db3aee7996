// poseidon.cuh — Poseidon permutation over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) and the
// sponge / compression modes plonky2's PoseidonHash uses.
//
// Replaces plonky2 0.2.2 hash/poseidon.rs (Poseidon::poseidon), hash/hashing.rs (hash_n_to_m_no_pad,
// compress) and Hasher::{hash_no_pad, hash_or_noop, two_to_one} for PoseidonHash — the hasher of
// `PoseidonGoldilocksConfig`, the only config the reference instantiates
// (city_rollup_core_worker/src/lib.rs:25-26; wrappers at city_crypto/src/hash/traits/hasher.rs:77-159).
//
// Schedule (one permutation per thread, state in registers; "v6", details at permute_nc below):
//   * the round-constant layer of round r+1 is folded into the MDS accumulators of round r;
//   * S-box x^7 = 2 squarings + 2 multiplications with non-canonical (u64) intermediates; the last product is
//     not reduced but handed to the MDS layer as two limbs;
//   * the MDS layer (circulant [17,15,41,16,2,28,13,13,39,18,34,20] + diag(8,0,..)) runs on the FP64 pipe
//     as exact integer arithmetic on the two 32-bit limbs of every lane (see mds_limbs_biased below), then one
//     integer fold per lane using 2^64 = 2^32 - 1 where a 64-bit integer is needed (S-box inputs), a 6-instruction
//     lazy fold where it is not.  (History, profiles/r01_summary.md: v1 IMAD.WIDE MDS saturated the FMA-heavy
//     pipe; v2 22/21/21-bit limbs on plain IMAD; v3 FP64 MDS; v6 = this.)
//   * nothing is canonicalised until the digest is written.
#pragma once
#include "gl64.cuh"

namespace poseidon {

// 30 x 12 round constants followed by 12 zeros (the "next round" constants of the last round)
__constant__ uint64_t RC[372] = {
#include "poseidon_rc.inc"
};

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
#ifdef P2B_SBOX_V3  // tuning builds only
  uint64_t x2 = gl::mul_nc(x, x);
  uint64_t x4 = gl::mul_nc(x2, x2);
  uint64_t x3 = gl::mul_nc(x, x2);
  return gl::mul_nc(x3, x4);
#else  // 14 wide multiplies instead of 20 (they share an execution unit with the DFMAs of the MDS layers, gl64.cuh)
  uint64_t x2 = gl::sqr_nc(x);
  uint64_t x4 = gl::sqr_nc(x2);
  uint64_t x3 = gl::mul_nc_lw(x, x2);
  return gl::mul_nc_lw(x3, x4);
#endif
}

// ---- MDS layer on the FP64 pipe -----------------------------------------------------------------------
// B200 keeps a full-rate FP64 pipe (measured 56-63 DFMA/clk/SM, dual-issuing with the integer pipes:
// profiles/int32_peak.json), and a DFMA multiplies a 32-bit limb by a small MDS entry and accumulates it
// exactly (every partial sum < 2^53).  Per 32-bit limb set the 12x12 circulant is split by
// x^12 - 1 = (x^6 - 1)(x^6 + 1): with p_k = s_k + s_{k+6}, m_k = s_k - s_{k+6},
//     out[r] = P[r] + M[r],  out[r+6] = P[r] - M[r],
//     P[r] = sum_k Dh[(k-r) mod 6] p_k            Dh = (C[j] + C[j+6]) / 2 = [15,14,40,17,18,24]
//     M[r] = sum_k +-Eh[(k-r) mod 6] m_k          Eh = (C[j] - C[j+6]) / 2 = [2,1,1,-1,-16,4]  (sign - on wrap)
// (all entries of C[j] +- C[j+6] are even), 72 instead of 144 multiply-adds; diag(8,0,..,0) adds 4 s_0 to P[0]
// and M[0].  Integer <-> double conversion is free of I2F/F2I: a limb x becomes the double 2^52 + x by
// pairing it with the high word 0x43300000, the biases cancel in m_k and are removed from p_k by one DADD,
// and the chain initialisers (poseidon_rc_f64.inc) carry 2^52 + the next round's constant, so that the low
// 42 mantissa bits of every result ARE the integer row sum.  208 FP64 instructions per layer replace
// 3 x 144 IMAD + limb splitting; the FMA-heavy pipe is left to the S-boxes.
__constant__ unsigned long long RCF[720] = {
#include "poseidon_rc_f64.inc"
};

#define P2B_TWO52_HI 0x43300000u

// a + b / a - b on the FP64 pipe (issuing them as DFMA with a unit multiplier was measured and is slower:
// profiles/r01_poseidon_v6_experiments.md)
__device__ __forceinline__ double dadd(double a, double b) { return a + b; }
__device__ __forceinline__ double dsub(double a, double b) { return a - b; }

// a + 2^32 b (mod p) for a, b < 2^51 given as the bit patterns of 2^52 + a and 2^52 + b
__device__ __forceinline__ uint64_t fold_f64(double ya, double yb) {
  uint32_t a_lo = (uint32_t)__double2loint(ya), a_hw = (uint32_t)__double2hiint(ya);
  uint32_t b_lo = (uint32_t)__double2loint(yb), b_hw = (uint32_t)__double2hiint(yb);
  uint32_t b_hi = b_hw - P2B_TWO52_HI;                   // bits >= 32 of b: weight 2^64 = 2^32 - 1
  uint32_t m = a_hw + b_hw - 2u * P2B_TWO52_HI;          // a_hi + b_hi (< 2^20)
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 yl,yh,c;\n\t"
      "sub.cc.u32 yl, 0, %4;\n\t"     // Y = (m << 32) - b_hi  (>= 0)
      "subc.u32 yh, %5, 0;\n\t"
      "add.cc.u32 yl, yl, %2;\n\t"    // X + Y, X = a_lo + 2^32 b_lo
      "addc.cc.u32 yh, yh, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, yl, c;\n\t"     // + c * (2^32 - 1)
      "subc.u32 yh, yh, 0;\n\t"
      "add.u32 %1, yh, c;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a_lo), "r"(b_lo), "r"(b_hi), "r"(m));
  return gl::pack(r0, r1);
}

// The same for b >= 2^32 (every v6 layer adds 2^32 to its hi row sums and takes 2^64 out of the constant: lift_hi in
// tools/gen_poseidon_v6_tables.py): -b_hi is then 2^32 - b_hi in 32 bits, its "no borrow" is the carry of a plain add,
// and a_hi + b_hi - 1 >= 0 — 8 instructions instead of 10.
__device__ __forceinline__ uint64_t fold_f64_b1(double ya, double yb) {
  uint32_t a_lo = (uint32_t)__double2loint(ya), a_hw = (uint32_t)__double2hiint(ya);
  uint32_t b_lo = (uint32_t)__double2loint(yb), b_hw = (uint32_t)__double2hiint(yb);
  uint32_t nb = P2B_TWO52_HI - b_hw;                          // 2^32 - b_hi
  uint32_t m1 = a_hw + b_hw - (2u * P2B_TWO52_HI + 1u);       // a_hi + b_hi - 1
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 lo,hi,c;\n\t"
      "add.cc.u32 lo, %2, %4;\n\t"     // a_lo - b_hi + 2^32 [no borrow]
      "addc.cc.u32 hi, %3, %5;\n\t"    // b_lo + a_hi + b_hi - [borrow]
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, lo, c;\n\t"      // + c * (2^32 - 1)
      "subc.u32 hi, hi, 0;\n\t"
      "add.u32 %1, hi, c;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a_lo), "r"(b_lo), "r"(nb), "r"(m1));
  return gl::pack(r0, r1);
}

// one limb set: b[k] = the double 2^52 + limb_k (limb_k < 2^42)  ->  y[r] = 2^52 + (MDS limb)[r] + K[r]
__device__ __forceinline__ void mds_limbs_biased(const double (&b)[12], const unsigned long long* __restrict__ init,
                                                 double (&y)[12]) {
  constexpr double Dh[6] = {15., 14., 40., 17., 18., 24.};
  constexpr double Eh[6] = {2., 1., 1., -1., -16., 4.};
  double p[6], m[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    m[k] = dsub(b[k], b[k + 6]);                              // the 2^52 biases cancel
    p[k] = dadd(b[k], dsub(b[k + 6], 9007199254740992.0));   // (2^52 + x_k) + (x_{k+6} - 2^52): both steps exact
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double P = __longlong_as_double((long long)init[2 * r]);
    double M = __longlong_as_double((long long)init[2 * r + 1]);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int j = (k - r + 12) % 12;  // circulant index of s_k in row r
      double d = Dh[j % 6], e = j < 6 ? Eh[j] : -Eh[j - 6];
      if (r == 0 && k == 0) d += 2., e += 2.;  // diag: + 4 x_0 = 2 p_0 + 2 m_0 on both chains
      P = fma(d, p[k], P);
      M = fma(e, m[k], M);
    }
    if (r == 0) {
      P = fma(2., m[0], P);
      M = fma(2., p[0], M);
    }
    y[r] = dadd(P, M);
    y[r + 6] = dsub(P, M);
  }
}

__device__ __forceinline__ double biased_lo(uint64_t v) { return __hiloint2double((int)P2B_TWO52_HI, (int)(uint32_t)v); }
__device__ __forceinline__ double biased_hi(uint64_t v) { return __hiloint2double((int)P2B_TWO52_HI, (int)(uint32_t)(v >> 32)); }

// s <- MDS * s + rc(next round); `init` = the 24 chain initialisers of that round
__device__ __forceinline__ void mds_layer(uint64_t (&s)[12], const unsigned long long* __restrict__ init) {
  double b[12], ylo[12], yhi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) b[i] = biased_lo(s[i]);
  mds_limbs_biased(b, init, ylo);
#pragma unroll
  for (int i = 0; i < 12; i++) b[i] = biased_hi(s[i]);
  mds_limbs_biased(b, init + 12, yhi);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = fold_f64(ylo[i], yhi[i]);
}

// Two consecutive partial rounds.  Only lane 0 goes through an S-box, so lanes 1..11 can stay in the FP64 domain
// between the two MDS layers: the first layer's row sums (2^52 + a 41-bit integer) ARE valid biased limbs for
// the second layer (its partial sums stay below 2^53: 264 * 2^41 + 2^52), and only lane 0 is folded to an
// integer, raised to the 7th power and converted back.  Saves 11 folds and 22 conversions per pair of rounds.
__device__ __forceinline__ void partial_round_pair(uint64_t (&s)[12], const unsigned long long* __restrict__ init) {
  double blo[12], bhi[12], ylo[12], yhi[12];
  s[0] = sbox7(s[0]);
#pragma unroll
  for (int i = 0; i < 12; i++) blo[i] = biased_lo(s[i]), bhi[i] = biased_hi(s[i]);
  mds_limbs_biased(blo, init, ylo);
  mds_limbs_biased(bhi, init + 12, yhi);
  const uint64_t s0 = sbox7(fold_f64(ylo[0], yhi[0]));
  ylo[0] = biased_lo(s0);
  yhi[0] = biased_hi(s0);
  mds_limbs_biased(ylo, init + 24, blo);
  mds_limbs_biased(yhi, init + 36, bhi);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = fold_f64(blo[i], bhi[i]);
}

// ---- v6 schedule --------------------------------------------------------------------------------------------
// Measured on B200 (profiles/r01_poseidon_v6_experiments.md): the permutation kernels run at ~0.65 warp
// instructions per cycle per scheduler whatever the occupancy and however the integer and FP64 streams are
// interleaved (S-box pairs fenced into their own basic blocks beside the previous pair's chain terms: no gain; wide
// multiplies traded for ALU instructions: 1 %), and time follows the instruction count.  v6 removes instructions:
//   * the two MDS layers of a partial-round pair are chained: the second layer takes the first layer's P / M chains
//     (p_k = 2 P_k, m_k = 2 M_k) instead of its outputs — 10 outputs, their bias and 15 p / m less per limb set;
//   * the last multiplication of every S-box is not reduced: the 128-bit product x3:x2:x1:x0 goes to the MDS layer as
//     the limbs lo = x0 - x2 - x3 + 2^33, hi = x1 + x2 (6 instructions instead of 12 + 2 moves);
//   * lanes 1..11 between two pairs are folded lazily to limbs of 33 bits (6 instructions instead of 10 + 2 moves);
//   * every layer's hi row sums are lifted by 2^32 (and 2^64 taken out of the constant), which shortens the 64-bit
//     fold to 8 instructions (fold_f64_b1);
//   * the P chains of a layer are split once more by x^6 - 1 = (x^3 - 1)(x^3 + 1) (p_rows_split: 31 instead of 37 FP64
//     instructions per limb set).
// The constant offsets that keep those limbs non-negative pass through the linear layer and are taken out of the chain
// initialisers (RC6, tools/gen_poseidon_v6_tables.py — which also checks this schedule operation by operation in
// exact integer arithmetic, incl. every FP64 bound, against the plain permutation).
__constant__ unsigned long long RC6[720] = {
#include "poseidon_rc_v6.inc"
};
#define P2B_LAZY_OFFSET 0x40000u  // OL = 2^18 (gen_poseidon_v6_tables.py)

__device__ __forceinline__ void pm_from_biased(double bk, double bk6, double& p, double& m) {
  m = dsub(bk, bk6);                               // the 2^52 biases cancel
  p = dadd(bk, dsub(bk6, 9007199254740992.0));     // (2^52 + x_k) + (x_{k+6} - 2^52): both steps exact
}
// ---- one MDS layer on one limb set, from p_k = x_k + x_{k+6} and m_k = x_k - x_{k+6} ------------------------------
// P[r] = sum_k Dh[(k - r) mod 6] p_k is a cyclic convolution of length 6 and splits once more by
// x^6 - 1 = (x^3 - 1)(x^3 + 1): with u_k = p_k + p_{k+3}, v_k = p_k - p_{k+3} (k < 3)
//     S[r] = sI[r] + 16 (u_0 + u_1 + u_2) + 16 u_{(r+2) mod 3}        ((Dh[j] + Dh[j+3]) / 2 = [16, 16, 32])
//     T[r] = tI[r] + sum_k Ht[(k - r) mod 6] v_k                      ((Dh[j] - Dh[j+3]) / 2 = [-1, -2, 8], antiperiodic)
//     P[r] = S[r] + T[r],   P[r + 3] = S[r] - T[r],   P[0] += 2 p_0 + 2 m_0 (the diagonal 8 x_0)
// 31 FP64 instructions instead of 37.  `init` = the layer's table row: slots 2 rr hold sI[0..2], tI[0..2]
// (= (P_init[r] +- P_init[r + 3]) / 2, tools/gen_poseidon_v6_tables.py), slots 2 rr + 1 the M initialisers.
// SCALED: ph[1..5] / mh[1..5] hold HALF of p_k / m_k (they are the previous layer's P / M chains, p_k = 2 P_k).
template <bool SCALED>
__device__ __forceinline__ void p_rows_split(double p0, double m0, const double (&ph)[6],
                                             const unsigned long long* __restrict__ init, double (&P)[6]) {
  constexpr double s = SCALED ? 2. : 1.;
  const double u0 = SCALED ? fma(2., ph[3], p0) : dadd(p0, ph[3]);
  const double v0 = SCALED ? fma(-2., ph[3], p0) : dsub(p0, ph[3]);
  const double u1 = dadd(ph[1], ph[4]), v1 = dsub(ph[1], ph[4]);
  const double u2 = dadd(ph[2], ph[5]), v2 = dsub(ph[2], ph[5]);
  const double u12 = dadd(u1, u2);
  const double U = SCALED ? fma(2., u12, u0) : dadd(u0, u12);
  double S0 = fma(16., U, __longlong_as_double((long long)init[0]));
  double S1 = fma(16., U, __longlong_as_double((long long)init[2]));
  double S2 = fma(16., U, __longlong_as_double((long long)init[4]));
  S0 = fma(16. * s, u2, S0);
  S1 = fma(16., u0, S1);
  S2 = fma(16. * s, u1, S2);
  double T0 = dsub(__longlong_as_double((long long)init[6]), v0);
  double T1 = fma(-8., v0, __longlong_as_double((long long)init[8]));
  double T2 = fma(2., v0, __longlong_as_double((long long)init[10]));
  T0 = fma(-2. * s, v1, T0);
  T1 = fma(-1. * s, v1, T1);
  T2 = fma(-8. * s, v1, T2);
  T0 = fma(8. * s, v2, T0);
  T1 = fma(-2. * s, v2, T1);
  T2 = fma(-1. * s, v2, T2);
  P[0] = dadd(S0, T0);
  P[3] = dsub(S0, T0);
  P[1] = dadd(S1, T1);
  P[4] = dsub(S1, T1);
  P[2] = dadd(S2, T2);
  P[5] = dsub(S2, T2);
  P[0] = fma(2., p0, P[0]);
  P[0] = fma(2., m0, P[0]);
}
// M[r] = M_init[r] + sum_k +-Eh[(k - r) mod 6] m_k (sign - on wrap), M[0] += 2 m_0 + 2 p_0
template <bool SCALED, int r>
__device__ __forceinline__ double m_row(double p0, double m0, const double (&mh)[6], const unsigned long long* __restrict__ init) {
  constexpr double Eh[6] = {2., 1., 1., -1., -16., 4.};
  constexpr double s = SCALED ? 2. : 1.;
  double M = __longlong_as_double((long long)init[2 * r + 1]);
#pragma unroll
  for (int k = 1; k < 6; k++) {
    const int j = (k - r + 12) % 12;
    const double e = j < 6 ? Eh[j % 6] : -Eh[j % 6];
    M = fma(e * s, mh[k], M);
  }
  {
    constexpr int j = (0 - r + 12) % 12;
    constexpr double e = (j < 6 ? Eh[j % 6] : -Eh[j % 6]) + (r == 0 ? 2. : 0.);
    M = fma(e, m0, M);
  }
  if (r == 0) M = fma(2., p0, M);
  return M;
}
template <bool SCALED>
__device__ __forceinline__ void m_rows(double p0, double m0, const double (&mh)[6], const unsigned long long* __restrict__ init,
                                       double (&M)[6]) {
  M[0] = m_row<SCALED, 0>(p0, m0, mh, init);
  M[1] = m_row<SCALED, 1>(p0, m0, mh, init);
  M[2] = m_row<SCALED, 2>(p0, m0, mh, init);
  M[3] = m_row<SCALED, 3>(p0, m0, mh, init);
  M[4] = m_row<SCALED, 4>(p0, m0, mh, init);
  M[5] = m_row<SCALED, 5>(p0, m0, mh, init);
}
// one limb set of a layer whose twelve inputs are biased limbs b -> biased row sums y
__device__ __forceinline__ void mds_limbs_v6(const double (&b)[12], const unsigned long long* __restrict__ init, double (&y)[12]) {
  double p[6], m[6], P[6], M[6];
#pragma unroll
  for (int k = 0; k < 6; k++) pm_from_biased(b[k], b[k + 6], p[k], m[k]);
  p_rows_split<false>(p[0], m[0], p, init, P);
  m_rows<false>(p[0], m[0], m, init, M);
#pragma unroll
  for (int r = 0; r < 6; r++) {
    y[r] = dadd(P[r], M[r]);
    y[r + 6] = dsub(P[r], M[r]);
  }
}

// x^7 in limb form: blo = 2^52 + (x0 - x2 - x3 + 2^33), bhi = 2^52 + (x1 + x2) for the unreduced last product
__device__ __forceinline__ void sbox7_limbs(uint64_t x, double& blo, double& bhi) {
#ifdef P2B_SBOX_V3  // tuning builds only
  const uint64_t x2 = gl::mul_nc(x, x);
  const uint64_t x4 = gl::mul_nc(x2, x2);
  const uint64_t x3 = gl::mul_nc(x, x2);
#else
  const uint64_t x2 = gl::sqr_nc(x);
  const uint64_t x4 = gl::sqr_nc(x2);
  const uint64_t x3 = gl::mul_nc_lw(x, x2);
#endif
  asm("{\n\t"
      ".reg .u32 x0,x1,x2,x3,l0,l1,h0,h1;\n\t"
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
      "sub.cc.u32 l0, x0, x2;\n\t"           // (2 : x0) - x2 - x3, the 2 sitting on top of the exponent word
      "subc.u32 l1, 0x43300002, 0;\n\t"
      "sub.cc.u32 l0, l0, x3;\n\t"
      "subc.u32 l1, l1, 0;\n\t"
      "add.cc.u32 h0, x1, x2;\n\t"
      "addc.u32 h1, 0x43300000, 0;\n\t"
      "mov.b64 %0, {l0, l1};\n\t"
      "mov.b64 %1, {h0, h1};\n\t"
      "}"
      : "=d"(blo), "=d"(bhi)
      : "r"((uint32_t)x3), "r"((uint32_t)(x3 >> 32)), "r"((uint32_t)x4), "r"((uint32_t)(x4 >> 32)));
}
// row sums a, b (as 2^52 + a, 2^52 + b; a, b < 2^50) -> limbs lo = a_lo - b_hi + 2^18, hi = b_lo + a_hi + b_hi
__device__ __forceinline__ void lazy_fold(double ya, double yb, double& blo, double& bhi) {
  asm("{\n\t"
      ".reg .u32 al,ah,bl,bh,u,t,l1,h1;\n\t"
      "mov.b64 {al, ah}, %2;\n\t"
      "mov.b64 {bl, bh}, %3;\n\t"
      "sub.u32 u, 0x43340000, bh;\n\t"       // 2^18 - b_hi
      "add.cc.u32 al, al, u;\n\t"
      "addc.u32 l1, 0x43300000, 0;\n\t"
      "add.u32 t, ah, bh;\n\t"
      "add.u32 t, t, 0x79a00000;\n\t"        // - 2 * 0x43300000 (mod 2^32)
      "add.cc.u32 bl, bl, t;\n\t"
      "addc.u32 h1, 0x43300000, 0;\n\t"
      "mov.b64 %0, {al, l1};\n\t"
      "mov.b64 %1, {bl, h1};\n\t"
      "}"
      : "=d"(blo), "=d"(bhi)
      : "d"(ya), "d"(yb));
}
// a 64-bit lane -> the limbs a lazy fold would have produced (lo + 2^18, hi)
__device__ __forceinline__ void limbs_from_u64(uint64_t v, double& blo, double& bhi) {
  asm("{\n\t"
      ".reg .u32 l0,l1;\n\t"
      "add.cc.u32 l0, %2, 0x40000;\n\t"
      "addc.u32 l1, 0x43300000, 0;\n\t"
      "mov.b64 %0, {l0, l1};\n\t"
      "mov.b64 %1, {%3, 0x43300000};\n\t"
      "}"
      : "=d"(blo), "=d"(bhi)
      : "r"((uint32_t)v), "r"((uint32_t)(v >> 32)));
}

// One full round: s <- MDS * sbox(s) + rc(next round)
__device__ __forceinline__ void full_round_v6(uint64_t (&s)[12], const unsigned long long* __restrict__ init) {
  double blo[12], bhi[12], ylo[12], yhi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) sbox7_limbs(s[i], blo[i], bhi[i]);
  mds_limbs_v6(blo, init, ylo);
  mds_limbs_v6(bhi, init + 12, yhi);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = fold_f64_b1(ylo[i], yhi[i]);
}

// Two consecutive partial rounds r, r + 1.  s0 = lane 0 (a 64-bit integer: it passes the S-boxes), zlo / zhi[1..11] =
// the other lanes in limb form; all updated in place.
// `vz` = lane_varying_zero(): with a warp-uniform table index ptxas loads the initialisers into UNIFORM registers, a
// DFMA takes only one non-register operand, and every chain then starts with two moves that put its multiplier (an
// immediate otherwise) into a register pair — 65 extra instructions per pair.
__device__ __forceinline__ uint32_t lane_varying_zero() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));  // < 2^31 on every lane
  return m >> 31;
}
// `mid` maps the computed S-box input of the second round to the value that is raised to the 7th power (identity for the
// permutation; the PoseidonGate evaluator records the difference to the wire and continues with the wire).
template <class MidHook>
__device__ __forceinline__ void partial_round_pair_v6_hook(uint64_t& s0, double (&zlo)[12], double (&zhi)[12], int r,
                                                           uint32_t vz, MidHook&& mid) {
  const unsigned long long* __restrict__ initA = RC6 + 24 * (r + vz);
  const unsigned long long* __restrict__ initB = initA + 24;
  double Pl[6], Ml[6], Ph[6], Mh[6];  // first layer
  double Ql[6], Nl[6], Qh[6], Nh[6];  // second layer
  double b0l, b0h, p0l, m0l, p0h, m0h;
  // ---- first layer: lane 0 from its S-box, lanes 1..11 from their limbs
  sbox7_limbs(s0, b0l, b0h);
  {
    double p[6], m[6];
#pragma unroll
    for (int k = 1; k < 6; k++) pm_from_biased(zlo[k], zlo[k + 6], p[k], m[k]);
    pm_from_biased(b0l, zlo[6], p0l, m0l);
    p[0] = m[0] = 0.;
    p_rows_split<false>(p0l, m0l, p, initA, Pl);
    m_rows<false>(p0l, m0l, m, initA, Ml);
#pragma unroll
    for (int k = 1; k < 6; k++) pm_from_biased(zhi[k], zhi[k + 6], p[k], m[k]);
    pm_from_biased(b0h, zhi[6], p0h, m0h);
    p_rows_split<false>(p0h, m0h, p, initA + 12, Ph);
    m_rows<false>(p0h, m0h, m, initA + 12, Mh);
  }
  const double y0l = dadd(Pl[0], Ml[0]), y0h = dadd(Ph[0], Mh[0]);  // lanes 0 and 6 (biased): the only outputs formed
  const double y6l = dsub(Pl[0], Ml[0]), y6h = dsub(Ph[0], Mh[0]);
  // ---- second layer: the row pairs 1..5 of the first layer enter as its chains (halves of p_k / m_k)
  sbox7_limbs(mid(fold_f64_b1(y0l, y0h)), b0l, b0h);
  pm_from_biased(b0l, y6l, p0l, m0l);
  pm_from_biased(b0h, y6h, p0h, m0h);
  p_rows_split<true>(p0l, m0l, Pl, initB, Ql);
  m_rows<true>(p0l, m0l, Ml, initB, Nl);
  p_rows_split<true>(p0h, m0h, Ph, initB + 12, Qh);
  m_rows<true>(p0h, m0h, Mh, initB + 12, Nh);
  s0 = fold_f64_b1(dadd(Ql[0], Nl[0]), dadd(Qh[0], Nh[0]));
  lazy_fold(dsub(Ql[0], Nl[0]), dsub(Qh[0], Nh[0]), zlo[6], zhi[6]);
#pragma unroll
  for (int i = 1; i < 6; i++) {
    lazy_fold(dadd(Ql[i], Nl[i]), dadd(Qh[i], Nh[i]), zlo[i], zhi[i]);
    lazy_fold(dsub(Ql[i], Nl[i]), dsub(Qh[i], Nh[i]), zlo[i + 6], zhi[i + 6]);
  }
}

// ---- Q schedule: the two MDS layers of a partial-round pair as ONE application of M^2 ---------------------------
// With t = the state after the S-box of round r, u = M t + c_A, x2 = u_0 and y2 = x2^7 (or whatever `mid` puts in its
// place), the state after round r + 1 is  M (u + e_0 (y2 - u_0)) + c_B = M^2 t + col_0(M) (y2 - u_0) + (M c_A + c_B),
// and with M = C + 8 e_0 e_0^T (C the circulant), limb set by limb set:
//     out_i = (C^2 t)_i + C[i][0] w + 8 [i == 0] y2 + const_i,        w = 8 t_0 + y2 - X2,  X2 = row 0 of the first layer.
// C^2 is the circulant of C * C (row sum 2^16: 33-bit limbs stay below 2^50) and splits exactly like C:
//     D2 = [5252, 5904, 5072, 4988, 6384, 5168], E2 = [54, -72, -486, 252, -162, -36],
//     (D2[j] + D2[j+3]) / 2 = [5120, 6144, 5120], (D2[j] - D2[j+3]) / 2 = [132, -240, -48];
// C[i][0] w enters the P / M chains with the coefficients lane 0 has in a plain C layer.  Rows 1..5 of the first layer are
// never formed: 127 FP64 instructions per limb set and pair instead of 171 (the permutation kernels are bound by the
// issue path FP64 shares with the wide integer multiplies).  Tables and the exact-arithmetic model:
// tools/gen_poseidon_v6_tables.py (permute_q; every intermediate checked against 2^53, limbs at the corners of their ranges).
__constant__ unsigned long long RCQ[308] = {
#include "poseidon_rc_q.inc"
};

struct QLimbState {
  double p[6], m[6], t0u, y0;
};
// first half: p / m of the twelve inputs, X2 = row 0 of the first layer (biased: 2^52 + row sum), t_0 without its bias
__device__ __forceinline__ void q_first(double b0, const double (&z)[12], const unsigned long long* __restrict__ init,
                                        QLimbState& q) {
  constexpr double Dh[6] = {15., 14., 40., 17., 18., 24.};
  constexpr double Eh[6] = {2., 1., 1., -1., -16., 4.};
  pm_from_biased(b0, z[6], q.p[0], q.m[0]);
#pragma unroll
  for (int k = 1; k < 6; k++) pm_from_biased(z[k], z[k + 6], q.p[k], q.m[k]);
  double P = __longlong_as_double((long long)init[0]);
  double M = __longlong_as_double((long long)init[1]);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    P = fma(Dh[k] + (k == 0 ? 2. : 0.), q.p[k], P);
    M = fma(Eh[k] + (k == 0 ? 2. : 0.), q.m[k], M);
  }
  P = fma(2., q.m[0], P);  // the diagonal 8 t_0 = 4 p_0 + 4 m_0, half on each chain
  M = fma(2., q.p[0], M);
  q.y0 = dadd(P, M);
  q.t0u = dsub(b0, 4503599627370496.0);
}
// second half: b0 = the biased limb of y2 -> the twelve biased row sums of the pair
__device__ __forceinline__ void q_second(double b0, const QLimbState& q, const unsigned long long* __restrict__ init,
                                         double (&y)[12]) {
  constexpr double E2[6] = {54., -72., -486., 252., -162., -36.};
  constexpr double H2[6] = {132., -240., -48., -132., 240., 48.};
  constexpr double Dh[6] = {15., 14., 40., 17., 18., 24.};
  constexpr double Eh[6] = {2., 1., 1., -1., -16., 4.};
  const double d = dsub(b0, q.y0);               // y2 - X2: the 2^52 biases cancel
  const double w = fma(8., q.t0u, d);
  const double y2u = dsub(b0, 4503599627370496.0);
  const double u0 = dadd(q.p[0], q.p[3]), v0 = dsub(q.p[0], q.p[3]);
  const double u1 = dadd(q.p[1], q.p[4]), v1 = dsub(q.p[1], q.p[4]);
  const double u2 = dadd(q.p[2], q.p[5]), v2 = dsub(q.p[2], q.p[5]);
  const double U = dadd(dadd(u0, u1), u2);
  const double uu[3] = {u0, u1, u2}, vv[3] = {v0, v1, v2};
  double P[6], M[6];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    double S = fma(5120., U, __longlong_as_double((long long)init[2 + r]));
    S = fma(1024., uu[(r + 1) % 3], S);
    double T = __longlong_as_double((long long)init[5 + r]);
#pragma unroll
    for (int k = 0; k < 3; k++) T = fma(H2[(k - r + 6) % 6], vv[k], T);
    P[r] = dadd(S, T);
    P[r + 3] = dsub(S, T);
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double Mr = __longlong_as_double((long long)init[8 + r]);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int j = (k - r + 12) % 12;
      Mr = fma(j < 6 ? E2[j] : -E2[j - 6], q.m[k], Mr);
    }
    // the rank-1 term C[i][0] w: lane 0's coefficients in a plain C layer
    const int j0 = (12 - r) % 12;
    P[r] = fma(Dh[j0 % 6], w, P[r]);
    M[r] = fma(j0 < 6 ? Eh[j0] : -Eh[j0 - 6], w, Mr);
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
    y[r] = dadd(P[r], M[r]);
    y[r + 6] = dsub(P[r], M[r]);
  }
  y[0] = fma(8., y2u, y[0]);
}

// Two consecutive partial rounds r, r + 1 (r = 4, 6, .., 24), same contract as partial_round_pair_v6_hook.
template <class MidHook>
__device__ __forceinline__ void partial_round_pair_q_hook(uint64_t& s0, double (&zlo)[12], double (&zhi)[12], int r,
                                                          uint32_t vz, MidHook&& mid) {
  const unsigned long long* __restrict__ init = RCQ + 28 * (((r - 4) >> 1) + vz);
  double b0l, b0h;
  QLimbState ql, qh;
  sbox7_limbs(s0, b0l, b0h);
  q_first(b0l, zlo, init, ql);
  q_first(b0h, zhi, init + 14, qh);
  sbox7_limbs(mid(fold_f64_b1(ql.y0, qh.y0)), b0l, b0h);
  double ylo[12], yhi[12];
  q_second(b0l, ql, init, ylo);
  q_second(b0h, qh, init + 14, yhi);
  s0 = fold_f64_b1(ylo[0], yhi[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) lazy_fold(ylo[i], yhi[i], zlo[i], zhi[i]);
}

#ifdef P2B_POSEIDON_PAIR_AB  // the round-1 pair (two chained layers), for A/B measurements
#define P2B_PAIR_HOOK partial_round_pair_v6_hook
#else
#define P2B_PAIR_HOOK partial_round_pair_q_hook
#endif
__device__ __forceinline__ void partial_round_pair_v6(uint64_t& s0, double (&zlo)[12], double (&zhi)[12], int r,
                                                      uint32_t vz) {
  P2B_PAIR_HOOK(s0, zlo, zhi, r, vz, [](uint64_t x) { return x; });
}

// In-place permutation.  Inputs: any u64.  Outputs: u64 congruent mod p (NOT canonical).
// `abort_after_first_half()` is evaluated once, after the first four full rounds (about a quarter of the work); when
// it returns true the permutation is abandoned (state undefined) and false is returned — the proof-of-work search
// drops candidates this way that another thread's hit has made irrelevant in the meantime.
template <class Abort>
__device__ __forceinline__ bool permute_nc_abortable(uint64_t (&s)[12], Abort&& abort_after_first_half) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_nc(s[i], RC[i]);  // RC entries are canonical
#ifdef P2B_POSEIDON_V3  // the previous schedule (tuning builds: tools/poseidon_bench.cu)
#pragma unroll 1
  for (int r = 0; r < 30;) {
    if (r < 4 || r >= 26) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
      mds_layer(s, RCF + 24 * r);
      r += 1;
    } else {
      partial_round_pair(s, RCF + 24 * r);  // the 22 partial rounds, two at a time
      r += 2;
    }
  }
#else
  // one copy of each loop body: the full rounds of both ends share theirs through the outer loop
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int i = 0; i < 4; i++) full_round_v6(s, RC6 + 24 * (26 * half + i));
    if (half == 0) {
      if (abort_after_first_half()) return false;
      double zlo[12], zhi[12];
      zlo[0] = zhi[0] = 0.;
#pragma unroll
      for (int i = 1; i < 12; i++) limbs_from_u64(s[i], zlo[i], zhi[i]);
      uint64_t s0 = s[0];
      const uint32_t vz = lane_varying_zero();
#pragma unroll 1
      for (int r = 4; r < 26; r += 2) partial_round_pair_v6(s0, zlo, zhi, r, vz);
      s[0] = s0;
#pragma unroll
      for (int i = 1; i < 12; i++) s[i] = gl::sub_nc(fold_f64(zlo[i], zhi[i]), (uint64_t)P2B_LAZY_OFFSET);
    }
  }
#endif
  return true;
}
__device__ __forceinline__ void permute_nc(uint64_t (&s)[12]) {
  permute_nc_abortable(s, [] { return false; });
}

__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
  permute_nc(s);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::canon(s[i]);
}

// ---- warp-cooperative permutation (latency path) --------------------------------------------------------
// A single permutation on one thread is a ~23k-instruction dependent stream (25-30 us): fine when a kernel has
// 10^5 independent permutations in flight, ruinous for the transcript (one sponge), the top levels of a Merkle
// tree and the small FRI layers, which are chains of a handful of permutations.  Here a warp shares one
// permutation: lane l < 12 holds state element l, the S-boxes of a full round run in parallel, and the MDS
// row of lane l gathers the other lanes with shuffles (the circulant makes the multiplier of step i the same
// on every lane).  21 of the 22 partial rounds are linearised over all 32 lanes (coop_partial_rounds21 below).
__device__ const uint64_t RC_G[372] = {  // the round constants again, in global memory: per-lane indexed loads
#include "poseidon_rc.inc"
};

__device__ __forceinline__ uint64_t coop_mds(uint64_t s, uint32_t l) {
  constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  // this path is latency bound (one warp, one dependent chain): the 12 products are independent and summed
  // as a tree (4 partial sums) instead of one 12-deep multiply-add chain
  uint64_t lo4[4] = {0, 0, 0, 0}, hi4[4] = {0, 0, 0, 0};  // sums over the low / high 32-bit halves, < 2^42
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t src = l + i;
    if (src >= 12) src -= 12;
    const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)s, src, 16);
    const uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(s >> 32), src, 16);
    const uint32_t c = C[i] + ((i == 0 && l == 0) ? 8u : 0u);  // diag(8, 0, ..., 0)
    lo4[i & 3] += (uint64_t)lo * c;
    hi4[i & 3] += (uint64_t)hi * c;
  }
  const uint64_t acc_lo = (lo4[0] + lo4[1]) + (lo4[2] + lo4[3]);
  const uint64_t acc_hi = (hi4[0] + hi4[1]) + (hi4[2] + hi4[3]);
  // acc_lo + 2^32 acc_hi mod p with 2^64 = 2^32 - 1
  const uint32_t b_hi = (uint32_t)(acc_hi >> 32);
  const uint64_t x = (acc_lo & 0xFFFFFFFFull) | (acc_hi << 32);
  const uint64_t y = (((acc_lo >> 32) + b_hi) << 32) - b_hi;
  uint64_t r = x + y;
  if (r < x) r += GL_EPS;  // wrapped: r < 2^43, no second carry
  return r;
}

// The first 21 of the 22 partial rounds as ONE dependent chain of S-boxes (tables and derivation:
// tools/gen_poseidon_partial_linear.py).  Only lane 0 passes an S-box in a partial round, so the S-box input x_r of
// partial round r and the state after any number of them are affine in the entering state s and the earlier S-box
// outputs:   x_r = X[r].s + XC[r] + sum_{k<r} C[r][k] sb_k,     sb_r = x_r^7.
// x_0 is s_0 plus a constant; the other 32 forms (x_1..x_20, and the 12 state elements after the 21st round) take one
// accumulator in each of the 32 lanes.  A round is "every lane raises the broadcast x_r to the 7th power and adds
// coefficient * sb_r to its accumulator, lane r broadcasts x_{r+1}": ~110 instructions instead of the ~190 of
// S-box + cross-lane MDS layer — and a lone warp is bound by its instruction count (one issue per ~2.7 cycles,
// profiles/r01_coop_permutation_latency_v2.txt), not by the dependency chain.  The 22nd partial round and the full
// rounds run the ordinary way on lanes 0..11.
#include "poseidon_partial_lin.inc"

__device__ __forceinline__ uint64_t shfl64(uint64_t v, uint32_t src) {
  const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src);
  const uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
  return gl::pack(lo, hi);
}

// s (lanes 0..11) = the state after the 4th full round's MDS layer (constants of round 4 not yet added); returns, in
// lanes 0..11, the state after the MDS layer of the 21st partial round (constants of round 25 not yet added)
__device__ __forceinline__ uint64_t coop_partial_rounds21(uint64_t s, uint32_t lane) {
  uint64_t sb = sbox7(gl::add_nc(shfl64(s, 0), PL_X0_CONST));  // independent of the accumulator set-up below
  uint64_t acc;
  {  // the s-dependent part of this lane's form: four independent multiply-add chains, not one 12-deep chain
    uint64_t q[4] = {PL_CONST[lane], 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 12; i++) q[i & 3] = gl::mad_nc(PL_INIT[i * 32 + lane], shfl64(s, i), q[i & 3]);
    acc = gl::add_nc(gl::add_nc(gl::canon(q[0]), q[1]), gl::canon(gl::add_nc(gl::canon(q[2]), q[3])));
  }
  uint64_t c = PL_SB[lane];
#pragma unroll 1
  for (int r = 0; r < 21; r++) {
    const uint64_t cn = PL_SB[(r < 20 ? r + 1 : 20) * 32 + lane];  // next round's coefficient: off the critical path
    acc = gl::mad_nc(c, sb, acc);
    if (r < 20) sb = sbox7(shfl64(acc, r));  // lane r now holds x_{r+1}
    c = cn;
  }
  return shfl64(acc, 20 + (lane < 12 ? lane : 0));
}

// One permutation shared by a warp.  lane = threadIdx & 31; lanes 0..11 hold the state elements, lanes 12..31 pass
// anything (they only work in the linearised partial rounds) and get garbage back; every lane of the warp must call
// (full-mask shuffles).
__device__ __forceinline__ uint64_t coop_permute_nc(uint64_t s, uint32_t lane) {
  const uint32_t l = lane & 15;  // the full rounds run in 16-lane shuffle segments; the upper one computes garbage
  const uint32_t lc = l < 12 ? l : 0;
  uint64_t rc = RC_G[lc];
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    s = gl::add_nc(s, rc);
    rc = RC_G[12 * (r + 1) + lc];  // next round's constant: off the critical path
    s = coop_mds(sbox7(s), l);
  }
  s = coop_partial_rounds21(s, lane);
  // partial round 22 (round 25), then the last four full rounds
  s = gl::add_nc(s, RC_G[12 * 25 + lc]);
  if (l == 0) s = sbox7(s);
  s = coop_mds(s, l);
  rc = RC_G[12 * 26 + lc];
#pragma unroll 1
  for (int r = 26; r < 30; r++) {
    s = gl::add_nc(s, rc);
    rc = RC_G[12 * (r + 1) + lc];  // row 30 is zero padding
    s = coop_mds(sbox7(s), l);
  }
  return s;
}

// two_to_one(l, r) = permute([l, r, 0,0,0,0])[0..4]
__device__ __forceinline__ void two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
  permute_nc(s);
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
}

}  // namespace poseidon
