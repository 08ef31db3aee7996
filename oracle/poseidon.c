/* poseidon.c — Poseidon over Goldilocks, width 12, and the sponge/compression modes plonky2 uses.
 * TEST INFRASTRUCTURE (see p2oracle.h).
 * Restates plonky2 0.2.2 hash/poseidon.rs (naive schedule: constant layer, S-box layer, MDS layer per
 * round; 4 full + 22 partial + 4 full), hash/hashing.rs (hash_n_to_m_no_pad, compress) and
 * hash/poseidon.rs PoseidonHash::{hash_no_pad,two_to_one}, Hasher::hash_or_noop — none on disk
 * (SURVEY A.6).  The reference wraps exactly these at city_crypto/src/hash/traits/hasher.rs:77-159.
 * Pinned by city_crypto/src/hash/cached_zero_hashes.rs:10-1036 (two_to_one chain) and :1039-2066
 * (9-element hash_no_pad then two_to_one chain). */
#include "gl_inline.h"

static const uint64_t RC[360] = {
#include "poseidon_rc.inc"
};
static const uint64_t MDS_DIAG0 = 8;

static inline uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gli_mul(x, x), x4 = gli_mul(x2, x2), x3 = gli_mul(x, x2);
  return gli_mul(x3, x4);
}

/* MDS layer on the two 32-bit halves of every lane: the circulant entries are < 2^6, so the 12-term sums
 * of 32-bit halves stay below 2^42 in plain u64 accumulators (the same restructuring plonky2's optimised
 * Goldilocks MDS uses; the result equals the u128 row sums of the naive schedule). */
static inline void mds_layer(uint64_t s[12]) {
  uint32_t lo[24], hi[24];
  uint64_t al[12], ah[12];
  static const uint32_t C32[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = (uint32_t)s[i];
    hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32);
  }
  for (int r = 0; r < 12; r++) {
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 12; i++) {
      a += (uint64_t)lo[i + r] * C32[i];
      b += (uint64_t)hi[i + r] * C32[i];
    }
    al[r] = a;
    ah[r] = b;
  }
  al[0] += (uint64_t)lo[0] * MDS_DIAG0;
  ah[0] += (uint64_t)hi[0] * MDS_DIAG0;
  for (int r = 0; r < 12; r++) {
    /* al + 2^32 ah, ah = ah1:ah0 :  al + ah1*(2^32-1) + (ah0 << 32)  with one carry fold */
    uint64_t m = al[r] + (ah[r] >> 32) * GL_EPS, t;
    if (__builtin_add_overflow(m, ah[r] << 32, &t)) t += GL_EPS;
    s[r] = t >= GL_P ? t - GL_P : t;
  }
}

/* The literal round structure.  (Measured here: 4.7 us per permutation per core with gcc's AVX2 code for the
 * 32-bit-half MDS; the fast-partial-round form below is NOT faster in this scalar C — 5.3 us — so the hot loops
 * and the CPU baseline keep this one.) */
void poseidon_permute(uint64_t s[12]) {
  for (int i = 0; i < 12; i++) s[i] = s[i] >= GL_P ? s[i] - GL_P : s[i];
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], RC[12 * r + i]);  /* constants on ALL lanes */
    if (r < 4 || r >= 26) {
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    mds_layer(s);
  }
}

#include "poseidon_fast.inc"

/* The same permutation with plonky2's "fast partial rounds" (hash/poseidon.rs partial_first_constant_layer,
 * mds_partial_layer_init, mds_partial_layer_fast; tables derived by tools/gen_poseidon_fast_tables.py): the form
 * PoseidonGate constrains.  Cross-checked against poseidon_permute in tests/test_oracle_consistency.py. */
void poseidon_permute_fast(uint64_t s[12]) {
  for (int i = 0; i < 12; i++) s[i] = s[i] >= GL_P ? s[i] - GL_P : s[i];
  int r = 0;
  for (; r < 4; r++) {
    for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], RC[12 * r + i]);
    for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    mds_layer(s);
  }
  for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], PFAST_FIRST[i]);
  {
    uint64_t o[12];
    o[0] = s[0];
    for (int i = 0; i < 11; i++) {
      u128 acc = 0; /* 11 products < 2^128 / 2^4: fold the high part as we go */
      uint64_t hi_acc = 0;
      for (int j = 0; j < 11; j++) {
        u128 t = (u128)PFAST_INIT[i * 11 + j] * s[1 + j];
        acc += (uint64_t)t;
        hi_acc = gli_add(hi_acc, gli_reduce128((u128)(uint64_t)(t >> 64) << 64));
      }
      o[1 + i] = gli_add(gli_reduce128(acc), hi_acc);
    }
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
  for (int k = 0; k < 22; k++) {
    uint64_t s0 = sbox7(s[0]);
    if (k < 21) s0 = gli_add(s0, PFAST_POST[k]);
    uint64_t d = gli_mul(s0, 25);
    for (int i = 1; i < 12; i++) d = gli_add(d, gli_mul(s[i], PFAST_W_HATS[k * 11 + i - 1]));
    for (int i = 1; i < 12; i++) s[i] = gli_add(s[i], gli_mul(s0, PFAST_VS[k * 11 + i - 1]));
    s[0] = d;
  }
  for (r = 26; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], RC[12 * r + i]);
    for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    mds_layer(s);
  }
}

/* hash_n_to_m_no_pad with m = 4: overwrite-mode absorb of <=8 elements per permutation */
void poseidon_hash_no_pad(const uint64_t *in, size_t len, uint64_t out[4]) {
  uint64_t s[12] = {0};
  for (size_t off = 0; off < len; off += 8) {
    size_t k = len - off < 8 ? len - off : 8;
    for (size_t i = 0; i < k; i++) s[i] = in[off + i];
    poseidon_permute(s);
  }
  for (int i = 0; i < 4; i++) out[i] = gl_canon(s[i]);
}

/* Hasher::hash_or_noop: <= 4 elements are zero-padded into the digest, else hash_no_pad */
void poseidon_hash_or_noop(const uint64_t *in, size_t len, uint64_t out[4]) {
  if (len <= 4) {
    for (size_t i = 0; i < 4; i++) out[i] = i < len ? gl_canon(in[i]) : 0;
  } else {
    poseidon_hash_no_pad(in, len, out);
  }
}

void poseidon_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
  poseidon_permute(s);
  for (int i = 0; i < 4; i++) out[i] = s[i];
}
