/* gl_inline.h — inline Goldilocks helpers for the oracle's hot loops.  TEST INFRASTRUCTURE.
 * reduce128 follows the published plonky2_field 0.2.2 goldilocks_field.rs::reduce128:
 * 2^64 = 2^32 - 1 and 2^96 = -1 (mod p). */
#ifndef GL_INLINE_H
#define GL_INLINE_H
#include "p2oracle.h"
typedef unsigned __int128 u128;

static inline uint64_t gli_reduce128(u128 x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
  uint64_t t0, r;
  /* branch-free: the carry of t0 + t1 is a coin flip and would mispredict */
  uint64_t b = __builtin_sub_overflow(lo, hi_hi, &t0);
  t0 -= (0 - b) & GL_EPS;
  uint64_t t1 = hi_lo * GL_EPS;
  uint64_t c = __builtin_add_overflow(t0, t1, &r);
  r += (0 - c) & GL_EPS;
  return r - ((0 - (uint64_t)(r >= GL_P)) & GL_P);
}
static inline uint64_t gli_mul(uint64_t a, uint64_t b) { return gli_reduce128((u128)a * b); }
/* a, b canonical */
static inline uint64_t gli_add(uint64_t a, uint64_t b) {
  uint64_t s = a + b;
  return s - ((0 - (uint64_t)(s < a || s >= GL_P)) & GL_P);
}
static inline uint64_t gli_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline size_t bitrev(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
#endif
