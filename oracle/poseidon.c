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
