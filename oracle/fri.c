/* fri.c — Challenger, FRI commit phase, proof of work, and the verifier-side arity fold.
 * TEST INFRASTRUCTURE (see p2oracle.h).
 * Restates plonky2 0.2.2 iop/challenger.rs (Challenger: overwrite-mode duplex, pop from the end of
 * the squeezed rate), fri/prover.rs (fri_committed_trees, fri_proof_of_work) and fri/verifier.rs
 * (compute_evaluation) — none on disk.  Parameters follow
 * city_common_circuit/src/circuits/zk_signature2/mod.rs:38-50 (rate_bits 3, cap_height 4, pow 16,
 * arity bits [4,4]).  The fold convention is pinned by the stored proofs of qbench_data/example.bin
 * (tests/test_oracle_golden.py solves beta from one query round and checks the other 27). */
#include <stdlib.h>
#include <string.h>

#include "gl_inline.h"

void challenger_init(p2o_challenger *c) { memset(c, 0, sizeof(*c)); }

static void duplexing(p2o_challenger *c) {
  for (unsigned i = 0; i < c->n_in; i++) c->state[i] = c->in[i];
  c->n_in = 0;
  poseidon_permute(c->state);
  memcpy(c->out, c->state, 8 * sizeof(uint64_t));
  c->n_out = 8;
}

void challenger_observe(p2o_challenger *c, const uint64_t *elems, size_t n) {
  for (size_t i = 0; i < n; i++) {
    c->n_out = 0;
    c->in[c->n_in++] = gl_canon(elems[i]);
    if (c->n_in == 8) duplexing(c);
  }
}

uint64_t challenger_get(p2o_challenger *c) {
  if (c->n_in != 0 || c->n_out == 0) duplexing(c);
  return c->out[--c->n_out];
}

static void ext_add(const uint64_t a[2], const uint64_t b[2], uint64_t o[2]) {
  o[0] = gl_add(a[0], b[0]);
  o[1] = gl_add(a[1], b[1]);
}
static void ext_sub(const uint64_t a[2], const uint64_t b[2], uint64_t o[2]) {
  o[0] = gl_sub(a[0], b[0]);
  o[1] = gl_sub(a[1], b[1]);
}

void fri_committed_trees(const uint64_t *coeffs_in, const uint64_t *values_in, size_t len,
                         const unsigned *arity_bits, size_t n_layers, unsigned rate_bits,
                         unsigned cap_height, p2o_challenger *ch, uint64_t *caps_out,
                         uint64_t **layer_leaves_out, uint64_t **layer_digests_out,
                         uint64_t *final_poly_out, uint64_t *betas_out) {
  size_t n_cap = (size_t)1 << cap_height;
  uint64_t *coeffs = (uint64_t *)malloc(len * 16), *values = (uint64_t *)malloc(len * 16);
  memcpy(coeffs, coeffs_in, len * 16);
  memcpy(values, values_in, len * 16);
  uint64_t shift = 7;
  for (size_t l = 0; l < n_layers; l++) {
    unsigned ab = arity_bits[l];
    size_t arity = (size_t)1 << ab, n_leaves = len >> ab;
    unsigned log_len = 0;
    while (((size_t)1 << log_len) < len) log_len++;
    /* reverse_index_bits_in_place(values); leaves = chunks of `arity` ext values, flattened */
    uint64_t *leaves = (uint64_t *)malloc(len * 16);
    for (size_t j = 0; j < len; j++) {
      size_t r = bitrev(j, log_len);
      leaves[2 * j] = gl_canon(values[2 * r]);
      leaves[2 * j + 1] = gl_canon(values[2 * r + 1]);
    }
    uint64_t *digests = (uint64_t *)malloc(2 * (n_leaves - n_cap) * 32 + 32);
    uint64_t *cap = caps_out + l * n_cap * 4;
    merkle_tree_new(leaves, n_leaves, 2 * arity, cap_height, digests, cap);
    if (layer_leaves_out && layer_leaves_out[l]) memcpy(layer_leaves_out[l], leaves, len * 16);
    if (layer_digests_out && layer_digests_out[l])
      memcpy(layer_digests_out[l], digests, 2 * (n_leaves - n_cap) * 32);
    free(leaves);
    free(digests);
    challenger_observe(ch, cap, n_cap * 4); /* observe_cap */
    uint64_t beta[2];
    beta[0] = challenger_get(ch); /* get_extension_challenge = 2 base challenges in order */
    beta[1] = challenger_get(ch);
    if (betas_out) {
      betas_out[2 * l] = beta[0];
      betas_out[2 * l + 1] = beta[1];
    }
    /* coeffs <- chunks_exact(arity).map(reduce_with_powers(chunk, beta)) (Horner from the top) */
    for (size_t i = 0; i < n_leaves; i++) {
      uint64_t acc[2] = {0, 0};
      for (size_t j = arity; j-- > 0;) {
        uint64_t t[2];
        gl2_mul(acc, beta, t);
        ext_add(t, coeffs + 2 * (i * arity + j), acc);
      }
      coeffs[2 * i] = acc[0];
      coeffs[2 * i + 1] = acc[1];
    }
    len = n_leaves;
    shift = gl_pow(shift, arity);
    memcpy(values, coeffs, len * 16);
    gl2_coset_fft(values, log_len - ab, shift);
  }
  size_t n_final = len >> rate_bits; /* the truncated coefficients are zero for a valid codeword */
  for (size_t i = 0; i < 2 * n_final; i++) final_poly_out[i] = gl_canon(coeffs[i]);
  challenger_observe(ch, final_poly_out, 2 * n_final);
  free(coeffs);
  free(values);
}

static int pow_ok(const uint64_t base_state[12], unsigned pos, uint64_t w, unsigned pow_bits) {
  uint64_t s[12];
  memcpy(s, base_state, sizeof(s));
  s[pos] = w;
  poseidon_permute(s);
  /* squeeze().last() = state[7]; need pow_bits + (64 - 64) leading zeros of the canonical value */
  return (gl_canon(s[7]) >> (64 - pow_bits)) == 0;
}

int fri_pow_check(const p2o_challenger *ch, uint64_t w, unsigned pow_bits) {
  uint64_t s[12];
  memcpy(s, ch->state, sizeof(s));
  for (unsigned i = 0; i < ch->n_in; i++) s[i] = ch->in[i];
  if (ch->n_in == 8) return -1; /* cannot happen: a full buffer is always flushed */
  return pow_ok(s, ch->n_in, w, pow_bits);
}

uint64_t fri_proof_of_work(p2o_challenger *ch, unsigned pow_bits) {
  uint64_t s[12];
  memcpy(s, ch->state, sizeof(s));
  for (unsigned i = 0; i < ch->n_in; i++) s[i] = ch->in[i];
  unsigned pos = ch->n_in;
  uint64_t found = ~0ULL;
  for (uint64_t base = 0; found == ~0ULL; base += 4096) {
    uint64_t best = ~0ULL;
#pragma omp parallel for reduction(min : best)
    for (uint64_t w = base; w < base + 4096; w++)
      if (pow_ok(s, pos, w, pow_bits) && w < best) best = w;
    found = best;
  }
  challenger_observe(ch, &found, 1);
  (void)challenger_get(ch); /* pow_response */
  return found;
}

/* fri/verifier.rs::compute_evaluation: interpolate the arity points of the coset and evaluate at beta */
void fri_compute_evaluation(uint64_t x, unsigned x_index_within_coset, unsigned arity_bits,
                            const uint64_t *evals_in, const uint64_t beta[2], uint64_t out[2]) {
  size_t arity = (size_t)1 << arity_bits;
  uint64_t g = gl_root_of_unity(arity_bits);
  uint64_t *ev = (uint64_t *)malloc(arity * 16), *pts = (uint64_t *)malloc(arity * 8);
  for (size_t i = 0; i < arity; i++) { /* reverse_index_bits_in_place(evals) */
    size_t r = bitrev(i, arity_bits);
    ev[2 * i] = evals_in[2 * r];
    ev[2 * i + 1] = evals_in[2 * r + 1];
  }
  size_t rev = bitrev(x_index_within_coset, arity_bits);
  uint64_t coset_start = gl_mul(x, gl_pow(g, arity - rev));
  uint64_t y = 1;
  for (size_t i = 0; i < arity; i++) {
    pts[i] = gl_mul(coset_start, y);
    y = gl_mul(y, g);
  }
  /* Lagrange interpolation evaluated at beta (the barycentric form of interpolate()) */
  uint64_t acc[2] = {0, 0};
  for (size_t i = 0; i < arity; i++) {
    uint64_t num[2] = {1, 0}, den = 1;
    for (size_t j = 0; j < arity; j++) {
      if (j == i) continue;
      uint64_t d[2] = {gl_sub(beta[0], pts[j]), beta[1]}, t[2];
      gl2_mul(num, d, t);
      num[0] = t[0];
      num[1] = t[1];
      den = gl_mul(den, gl_sub(pts[i], pts[j]));
    }
    uint64_t di = gl_inv(den), term[2] = {gl_mul(num[0], di), gl_mul(num[1], di)}, t2[2];
    gl2_mul(term, ev + 2 * i, t2);
    ext_add(acc, t2, acc);
  }
  out[0] = acc[0];
  out[1] = acc[1];
  free(ev);
  free(pts);
  (void)ext_sub;
}
