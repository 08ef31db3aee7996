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
//   * the MDS layer (circulant [17,15,41,16,2,28,13,13,39,18,34,20] + diag(8,0,..)) runs on three
//     22/21/21-bit limbs of every lane: 3 x 144 full-rate 32-bit IMADs into 32-bit accumulators (< 2^31,
//     no carries), then one fold per lane using 2^64 = 2^32 - 1.  (Measured on B200: IMAD.WIDE.U32 and
//     IMAD.HI issue at half the IMAD rate, and the first version of this kernel ran the FMA-heavy pipe at
//     92% — profiles/r01_leaf_hash_v1.md.)
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

// Round constants pre-split into the three MDS limbs (bits [0,22), [22,43), [43,64)).
__constant__ uint4 RCL[372] = {
#include "poseidon_rc_limbs.inc"
};

// (x << K) + a in one ALU-pipe instruction (LEA); written in PTX so that the front end cannot fold the
// shift-adds back into multiplications by 17/18
template <int K>
__device__ __forceinline__ uint32_t shl_add(uint32_t x, uint32_t a) {
  uint32_t t;
  asm("shl.b32 %0, %1, %2;" : "=r"(t) : "r"(x), "n"(K));
  return t + a;
}

// a0 + 2^22 a1 + 2^43 a2 (a_k < 2^31) -> u64 congruent mod p, using 2^64 = 2^32 - 1
__device__ __forceinline__ uint64_t fold_limbs(uint32_t a0, uint32_t a1, uint32_t a2) {
  uint32_t u_hi = a2 >> 21;  // bits >= 64 of a2 * 2^43
  uint32_t u_lo = a2 << 11;  // bits 32..63 (as the high word)
  // m = a0 + (a1 << 22) + u_hi * (2^32 - 1)  < 2^53: no overflow
  uint64_t m = ((uint64_t)a1 << 22) + a0 + ((uint64_t)u_hi << 32) - u_hi;
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 ml,mh,c;\n\t"
      "mov.b64 {ml,mh}, %2;\n\t"
      "add.cc.u32 mh, mh, %3;\n\t"  // + u_lo * 2^32
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, ml, c;\n\t"   // + c * (2^32 - 1)
      "subc.u32 mh, mh, 0;\n\t"
      "add.u32 %1, mh, c;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(m), "r"(u_lo));
  return gl::pack(r0, r1);
}

// s <- MDS * s + rc (rc = limbs of the next round's constants).
// Every lane is cut into limbs of 22/21/21 bits so that the 12-term row sums with the 6-bit circulant
// entries (sum 256, +8 on the diagonal) plus the constant limb stay below 2^31: 3 x 144 plain 32-bit
// IMADs (full rate on the FMA-heavy pipe) instead of 2 x 144 half-rate IMAD.WIDE, no carries at all.
__device__ __forceinline__ void mds_layer(uint64_t (&s)[12], const uint4* __restrict__ rc) {
  uint32_t l0[12], l1[12], l2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t lo = (uint32_t)s[i], hi = (uint32_t)(s[i] >> 32);
    l0[i] = lo & 0x3FFFFFu;
    l1[i] = __funnelshift_r(lo, hi, 22) & 0x1FFFFFu;
    l2[i] = hi >> 11;
  }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    uint4 k = rc[r];
    uint32_t a0 = k.x, a1 = k.y, a2 = k.z;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const uint32_t x0 = l0[(i + r) % 12], x1 = l1[(i + r) % 12], x2 = l2[(i + r) % 12];
      // Pipe balancing: the S-boxes keep the FMA-heavy pipe busier than the ALU pipe, so the terms whose
      // multiplier is 2^a (16, 2) or 2^a + 2^b (17, 18) are done as shift-adds on the ALU pipe.
      if (i == 3) {  // x16
        a0 = shl_add<4>(x0, a0), a1 = shl_add<4>(x1, a1), a2 = shl_add<4>(x2, a2);
      } else if (i == 4) {  // x2
        a0 = shl_add<1>(x0, a0), a1 = shl_add<1>(x1, a1), a2 = shl_add<1>(x2, a2);
      } else if (i == 0 && r != 0) {  // x17
        a0 = shl_add<4>(x0, a0 + x0), a1 = shl_add<4>(x1, a1 + x1), a2 = shl_add<4>(x2, a2 + x2);
      } else if (i == 9) {  // x18
        a0 = shl_add<4>(x0, shl_add<1>(x0, a0)), a1 = shl_add<4>(x1, shl_add<1>(x1, a1)),
        a2 = shl_add<4>(x2, shl_add<1>(x2, a2));
      } else {
        uint32_t c = MDSC[(r == 0 && i == 0) ? 12 : i];  // diag(8,0,...,0) folded into [12]
        a0 += x0 * c;
        a1 += x1 * c;
        a2 += x2 * c;
      }
    }
    s[r] = fold_limbs(a0, a1, a2);
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
    mds_layer(s, RCL + 12 * (r + 1));
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
