// poseidon.cuh — Poseidon permutation over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) and the
// sponge / compression modes plonky2's PoseidonHash uses.
//
// Replaces plonky2 0.2.2 hash/poseidon.rs (Poseidon::poseidon), hash/hashing.rs (hash_n_to_m_no_pad,
// compress) and Hasher::{hash_no_pad, hash_or_noop, two_to_one} for PoseidonHash — the hasher of
// `PoseidonGoldilocksConfig`, the only config the reference instantiates
// (city_rollup_core_worker/src/lib.rs:25-26; wrappers at city_crypto/src/hash/traits/hasher.rs:77-159).
//
// Schedule (one permutation per thread, state in registers):
//   * the round-constant layer of round r+1 is folded into the MDS accumulators of round r;
//   * S-box x^7 = 4 Goldilocks multiplications with non-canonical (u64) intermediates;
//   * the MDS layer (circulant [17,15,41,16,2,28,13,13,39,18,34,20] + diag(8,0,..)) runs on the two
//     32-bit halves of every lane: 2 x 144 IMAD.WIDE.U32 by a small immediate into 64-bit accumulators
//     (< 2^41, no overflow), then one 96-bit -> 64-bit fold per lane using 2^64 = 2^32 - 1;
//   * nothing is canonicalised until the digest is written.
#pragma once
#include "gl64.cuh"

namespace poseidon {

// 30 x 12 round constants followed by 12 zeros (the "next round" constants of the last round)
__constant__ uint64_t RC[372] = {
#include "poseidon_rc.inc"
};

// MDS multipliers live in the constant bank (used as c[bank][off] operands of IMAD.WIDE.U32): as
// immediates ptxas strength-reduces x2/x16/... into 4-instruction shift+add sequences.
// [0..11] circulant first row, [12] = circ[0] + diag[0].
__constant__ uint32_t MDSC[13] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20, 25};

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gl::mul_nc(x, x);
  uint64_t x4 = gl::mul_nc(x2, x2);
  uint64_t x3 = gl::mul_nc(x, x2);
  return gl::mul_nc(x3, x4);
}

// acc_lo + 2^32 * acc_hi (both < 2^42) -> u64 congruent mod p
__device__ __forceinline__ uint64_t fold96(uint64_t al, uint64_t ah) {
  uint32_t ah0 = (uint32_t)ah, ah1 = (uint32_t)(ah >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u64 m;\n\t"
      ".reg .u32 ml,mh,c;\n\t"
      "mad.wide.u32 m, %4, 0xFFFFFFFF, %2;\n\t"  // al + ah1*(2^32-1)  (2^64 = 2^32-1), < 2^43
      "mov.b64 {ml,mh}, m;\n\t"
      "add.cc.u32 mh, mh, %3;\n\t"               // + ah0 * 2^32
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, ml, c;\n\t"                // + c*(2^32-1)
      "subc.u32 mh, mh, 0;\n\t"
      "add.u32 %1, mh, c;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(al), "r"(ah0), "r"(ah1));
  return gl::pack(r0, r1);
}

// s <- MDS * s + rc  (rc = next round's constants, or nullptr-equivalent zero when last)
template <bool WITH_RC>
__device__ __forceinline__ void mds_layer(uint64_t (&s)[12], const uint64_t* __restrict__ rc) {
  uint32_t lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = (uint32_t)s[i];
    hi[i] = (uint32_t)(s[i] >> 32);
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    uint64_t al, ah;
    if (WITH_RC) {
      uint64_t k = rc[r];
      al = (uint32_t)k;
      ah = k >> 32;
    } else {
      al = 0;
      ah = 0;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) {
      uint32_t c = MDSC[(r == 0 && i == 0) ? 12 : i];  // diag(8,0,...,0) folded into [12]
      // explicit mad.wide.u32: nvcc otherwise strength-reduces x16/x2 into shift+mask sequences and
      // carries a dead "hi*c" term through every accumulation
      asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(al) : "r"(lo[(i + r) % 12]), "r"(c));
      asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(ah) : "r"(hi[(i + r) % 12]), "r"(c));
    }
    s[r] = fold96(al, ah);
  }
}

// In-place permutation.  Inputs: any u64.  Outputs: u64 congruent mod p (NOT canonical).
__device__ __forceinline__ void permute_nc(uint64_t (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_nc(s[i], RC[i]);  // RC entries are canonical
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    if (r < 4 || r >= 26) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    mds_layer<true>(s, RC + 12 * (r + 1));
  }
}

__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
  permute_nc(s);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::canon(s[i]);
}

// two_to_one(l, r) = permute([l, r, 0,0,0,0])[0..4]
__device__ __forceinline__ void two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
  permute_nc(s);
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
}

}  // namespace poseidon
