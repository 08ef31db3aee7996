/* field.c — Goldilocks field, quadratic extension.  TEST INFRASTRUCTURE (see p2oracle.h).
 * Restates plonky2_field 0.2.2 goldilocks_field.rs / extension/quadratic.rs (not on disk).
 * Pins: generator 7 and k_is = 7^i (city_common_circuit/src/circuits/zk_signature2/mod.rs:58-138). */
#include "gl_inline.h"

uint64_t gl_canon(uint64_t a) { return a >= GL_P ? a - GL_P : a; }

uint64_t gl_add(uint64_t a, uint64_t b) {
  u128 s = (u128)gl_canon(a) + gl_canon(b);
  return s >= GL_P ? (uint64_t)(s - GL_P) : (uint64_t)s;
}
uint64_t gl_sub(uint64_t a, uint64_t b) {
  a = gl_canon(a);
  b = gl_canon(b);
  return a >= b ? a - b : a + (GL_P - b);
}
uint64_t gl_mul(uint64_t a, uint64_t b) { return gli_mul(a, b); }
/* slow cross-check used by tests */
uint64_t gl_mul_slow(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % GL_P); }

uint64_t gl_pow(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  a = gl_canon(a);
  while (e) {
    if (e & 1) r = gl_mul(r, a);
    a = gl_mul(a, a);
    e >>= 1;
  }
  return r;
}
uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }

/* POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32) = 1753635133440165772 (verified in tests) */
uint64_t gl_root_of_unity(unsigned log_n) {
  uint64_t g = gl_pow(7, (GL_P - 1) >> 32);
  for (unsigned i = log_n; i < 32; i++) g = gl_mul(g, g);
  return g;
}

void gl2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
  uint64_t c0 = gl_add(gl_mul(a[0], b[0]), gl_mul(7, gl_mul(a[1], b[1])));
  uint64_t c1 = gl_add(gl_mul(a[0], b[1]), gl_mul(a[1], b[0]));
  out[0] = c0;
  out[1] = c1;
}
/* (a0 + a1 X)^-1 = (a0 - a1 X) / (a0^2 - 7 a1^2) */
void gl2_inv(const uint64_t a[2], uint64_t out[2]) {
  uint64_t norm = gl_sub(gl_mul(a[0], a[0]), gl_mul(7, gl_mul(a[1], a[1])));
  uint64_t ni = gl_inv(norm);
  out[0] = gl_mul(a[0], ni);
  out[1] = gl_mul(gl_sub(0, a[1]), ni);
}
