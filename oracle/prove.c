/* prove.c — the whole prover behind `circuit_data.prove(pw)` on the CPU, from the filled witness onwards.
 * TEST INFRASTRUCTURE (see p2oracle.h): the word-for-word reference of p2b_prove and the CPU arm of bench.py.
 *
 * Restates plonky2 0.2.2 plonk/prover.rs::prove_with_partition_witness (wires commit, betas / gammas, Z and partial
 * products commit, alphas, quotient commit, zeta, OpeningSet::new, observe_openings), fri/oracle.rs
 * PolynomialBatch::prove_openings (alpha.reduce_polys_base, divide_by_linear, shift_poly, lde + coset_fft) and
 * fri/prover.rs fri_proof (commit phase, proof of work, query rounds) — none on disk; the
 * reference's call sites are the `circuit_data.prove(pw)` of every worker circuit, e.g.
 * city_common_circuit/src/proof_minifier/pm_core.rs:151 and
 * city_rollup_circuit/src/block_circuits/ops/l2_transfer/circuit.rs:234.  It is a composition of the primitives of
 * ntt.c / merkle.c / fri.c / plonk.c and writes the proof as the same flat u64 words p2b_prove writes
 * (include/p2b.h: ProofWithPublicInputs' field order without length prefixes), so the two are compared with memcmp.
 * tests/verifier_ref.py::oracle_prove is the independent Python composition of the same primitives; the two agree
 * word for word (tests/test_prove_oracle.py).
 * PARITY: the FRI conventions (openings order, the two batches, alpha bookkeeping, fold) are pinned on the ten
 * proofs stored in qbench_data/example.bin; the PLONK transcript order, Z / quotient values and the upstream gate
 * formulas are "parity unpinned" by any reference fixture. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gl_inline.h"

/* P2O_TRACE=1: wall clock of every phase of p2o_prove on stderr (development aid) */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define MARK(what)                                                           \
  do {                                                                       \
    if (trace) {                                                             \
      double t__ = now_s();                                                  \
      fprintf(stderr, "[p2o_prove] %-24s %8.2f ms\n", what, (t__ - t_prev) * 1e3); \
      t_prev = t__;                                                          \
    }                                                                        \
  } while (0)

typedef struct {
  size_t n_cols;
  unsigned log_n, rate_bits, cap_height;
  uint64_t *coeffs;  /* n_cols x n, column-major */
  uint64_t *leaves;  /* N x n_cols, leaf order */
  uint64_t *digests; /* plonky2 layout */
  uint64_t *cap;     /* 2^cap_height x 4 */
} batch_t;

struct p2o_prover_data {
  batch_t cs;        /* prover_data.constants_sigmas_commitment */
  uint64_t *sigmas;  /* prover_data.sigmas: num_routed x n values on H */
};

static void batch_free(batch_t *b) {
  free(b->coeffs);
  free(b->leaves);
  free(b->digests);
  free(b->cap);
  memset(b, 0, sizeof(*b));
}

static void batch_alloc(batch_t *b, size_t n_cols, unsigned log_n, unsigned rate_bits, unsigned cap_height) {
  size_t n = (size_t)1 << log_n, N = n << rate_bits, n_cap = (size_t)1 << cap_height;
  b->n_cols = n_cols;
  b->log_n = log_n;
  b->rate_bits = rate_bits;
  b->cap_height = cap_height;
  b->coeffs = (uint64_t *)malloc(n_cols * n * 8);
  b->leaves = (uint64_t *)malloc(N * n_cols * 8);
  b->digests = (uint64_t *)malloc(2 * (N - n_cap) * 32 + 32);
  b->cap = (uint64_t *)malloc(n_cap * 32);
}

/* cols: n_cols x n contiguous, column-major */
static void batch_values(batch_t *b, const uint64_t *vals, size_t n_cols, unsigned log_n, unsigned rate_bits,
                         unsigned cap_height) {
  size_t n = (size_t)1 << log_n;
  batch_alloc(b, n_cols, log_n, rate_bits, cap_height);
  const uint64_t **ptrs = (const uint64_t **)malloc(n_cols * sizeof(*ptrs));
  for (size_t c = 0; c < n_cols; c++) ptrs[c] = vals + c * n;
  batch_from_values(ptrs, n_cols, log_n, rate_bits, cap_height, b->coeffs, b->leaves, b->digests, b->cap);
  free(ptrs);
}

static void batch_coeffs(batch_t *b, const uint64_t *coeffs, size_t n_cols, unsigned log_n, unsigned rate_bits,
                         unsigned cap_height) {
  size_t n = (size_t)1 << log_n;
  batch_alloc(b, n_cols, log_n, rate_bits, cap_height);
  memcpy(b->coeffs, coeffs, n_cols * n * 8);
  const uint64_t **ptrs = (const uint64_t **)malloc(n_cols * sizeof(*ptrs));
  for (size_t c = 0; c < n_cols; c++) ptrs[c] = b->coeffs + c * n;
  batch_from_coeffs(ptrs, n_cols, log_n, rate_bits, cap_height, b->leaves, b->digests, b->cap);
  free(ptrs);
}

p2o_prover_data *p2o_prover_data_new(const p2o_circuit *c, const uint64_t *const *cs_cols, unsigned rate_bits,
                                     unsigned cap_height) {
  size_t n = (size_t)1 << c->degree_bits, w = (size_t)c->num_constants + c->num_routed_wires;
  p2o_prover_data *pd = (p2o_prover_data *)calloc(1, sizeof(*pd));
  uint64_t *vals = (uint64_t *)malloc(w * n * 8);
  for (size_t k = 0; k < w; k++) memcpy(vals + k * n, cs_cols[k], n * 8);
  batch_values(&pd->cs, vals, w, c->degree_bits, rate_bits, cap_height);
  pd->sigmas = (uint64_t *)malloc((size_t)c->num_routed_wires * n * 8);
  memcpy(pd->sigmas, vals + (size_t)c->num_constants * n, (size_t)c->num_routed_wires * n * 8);
  free(vals);
  return pd;
}
void p2o_prover_data_free(p2o_prover_data *pd) {
  if (!pd) return;
  batch_free(&pd->cs);
  free(pd->sigmas);
  free(pd);
}
void p2o_prover_data_cap(const p2o_prover_data *pd, uint64_t *out) {
  memcpy(out, pd->cs.cap, ((size_t)1 << pd->cs.cap_height) * 32);
}

static size_t fri_words(const size_t widths[4], const unsigned cap_heights[4], unsigned log_n, const p2o_fri_params *fp) {
  unsigned log_N = log_n + fp->rate_bits, lc = log_N;
  size_t per_query = 0;
  for (int o = 0; o < 4; o++) per_query += widths[o] + 4 * (size_t)(log_N - cap_heights[o]);
  for (unsigned l = 0; l < fp->n_layers; l++) {
    lc -= fp->reduction_arity_bits[l];
    per_query += ((size_t)2 << fp->reduction_arity_bits[l]) + 4 * (size_t)(lc - fp->cap_height);
  }
  size_t n_final = ((size_t)1 << lc) >> fp->rate_bits;
  return (size_t)fp->n_layers * ((size_t)4 << fp->cap_height) + (size_t)fp->num_query_rounds * per_query + 2 * n_final + 1;
}

size_t p2o_proof_len(const p2o_circuit *c, unsigned cs_cap_height, const p2o_fri_params *fp, size_t n_pis) {
  unsigned sum = 0;
  if (fp->n_layers > 16) return 0;
  for (unsigned l = 0; l < fp->n_layers; l++) sum += fp->reduction_arity_bits[l];
  if (sum > c->degree_bits) return 0;
  const size_t nch = c->num_challenges;
  size_t widths[4] = {(size_t)c->num_constants + c->num_routed_wires, c->num_wires, nch * (1 + c->num_partial_products),
                      nch * c->quotient_degree_factor};
  unsigned caps[4] = {cs_cap_height, fp->cap_height, fp->cap_height, fp->cap_height};
  size_t n_open = widths[0] + widths[1] + widths[2] + nch + widths[3];
  return 3 * ((size_t)4 << fp->cap_height) + 2 * n_open + fri_words(widths, caps, c->degree_bits, fp) + n_pis;
}

/* PolynomialCoeffs::to_extension().eval(z): Horner from the top coefficient */
static void eval_ext(const uint64_t *coeffs, size_t n, const uint64_t z[2], uint64_t out[2]) {
  uint64_t acc[2] = {0, 0};
  for (size_t k = n; k-- > 0;) {
    uint64_t t[2];
    gl2_mul(acc, z, t);
    acc[0] = gli_add(t[0], gl_canon(coeffs[k]));
    acc[1] = t[1];
  }
  out[0] = acc[0];
  out[1] = acc[1];
}

static void eval_batch(const batch_t *b, size_t first, size_t count, const uint64_t z[2], uint64_t *out) {
  size_t n = (size_t)1 << b->log_n;
#pragma omp parallel for schedule(dynamic)
  for (size_t i = 0; i < count; i++) eval_ext(b->coeffs + (first + i) * n, n, z, out + 2 * i);
}

size_t p2o_prove(const p2o_circuit *c, const p2o_prover_data *pd, const uint64_t *circuit_digest,
                 const uint64_t *const *wire_cols, const uint64_t *public_inputs, size_t n_pis,
                 const p2o_fri_params *fp, uint64_t *out, size_t out_cap) {
  const unsigned log_n = c->degree_bits, rb = fp->rate_bits, cap_h = fp->cap_height, log_N = log_n + rb;
  const size_t n = (size_t)1 << log_n, N = n << rb, nch = c->num_challenges, n_cap = (size_t)1 << cap_h;
  const size_t npp = c->num_partial_products, qdf = c->quotient_degree_factor;
  const size_t proof_len = p2o_proof_len(c, pd->cs.cap_height, fp, n_pis);
  if (proof_len == 0 || out_cap < proof_len || nch > 4) return 0;
  const int trace = getenv("P2O_TRACE") != NULL;
  double t_prev = now_s();
  const batch_t *cs = &pd->cs;
  batch_t wi, zs, qt;
  uint64_t pih[4];
  poseidon_hash_no_pad(public_inputs, n_pis, pih);
  /* wires commitment */
  uint64_t *wvals = (uint64_t *)malloc((size_t)c->num_wires * n * 8);
  for (size_t w = 0; w < c->num_wires; w++) memcpy(wvals + w * n, wire_cols[w], n * 8);
  batch_values(&wi, wvals, c->num_wires, log_n, rb, cap_h);
  MARK("wires commit");
  p2o_challenger ch;
  challenger_init(&ch);
  challenger_observe(&ch, circuit_digest, 4);
  challenger_observe(&ch, pih, 4);
  challenger_observe(&ch, wi.cap, 4 * n_cap);
  uint64_t betas[4], gammas[4], alphas[4];
  for (size_t i = 0; i < nch; i++) betas[i] = challenger_get(&ch);
  for (size_t i = 0; i < nch; i++) gammas[i] = challenger_get(&ch);
  /* Z and partial products */
  const size_t w_zs = nch * (1 + npp), w_q = nch * qdf;
  uint64_t *zvals = (uint64_t *)malloc(w_zs * n * 8);
  plonk_partial_products_and_zs(c, wvals, pd->sigmas, betas, gammas, zvals);
  MARK("partial products / Z");
  batch_values(&zs, zvals, w_zs, log_n, rb, cap_h);
  free(zvals);
  free(wvals);
  MARK("zs commit");
  challenger_observe(&ch, zs.cap, 4 * n_cap);
  for (size_t i = 0; i < nch; i++) alphas[i] = challenger_get(&ch);
  /* quotient */
  uint64_t *chunks = (uint64_t *)malloc(w_q * n * 8);
  plonk_compute_quotient_polys(c, rb, cs->leaves, wi.leaves, zs.leaves, pih, betas, gammas, alphas, chunks);
  MARK("compute_quotient_polys");
  batch_coeffs(&qt, chunks, w_q, log_n, rb, cap_h);
  free(chunks);
  MARK("quotient commit");
  challenger_observe(&ch, qt.cap, 4 * n_cap);
  uint64_t zeta[2], zeta_next[2];
  zeta[0] = challenger_get(&ch);
  zeta[1] = challenger_get(&ch);
  const uint64_t g = gl_root_of_unity(log_n);
  zeta_next[0] = gli_mul(zeta[0], g);
  zeta_next[1] = gli_mul(zeta[1], g);

  /* proof words: caps | openings | FRI proof | public inputs */
  size_t off = 0;
  memcpy(out + off, wi.cap, n_cap * 32), off += 4 * n_cap;
  memcpy(out + off, zs.cap, n_cap * 32), off += 4 * n_cap;
  memcpy(out + off, qt.cap, n_cap * 32), off += 4 * n_cap;
  uint64_t *o_constants = out + off; /* constants | sigmas = all of constants_sigmas */
  uint64_t *o_wires = o_constants + 2 * cs->n_cols;
  uint64_t *o_zs = o_wires + 2 * (size_t)c->num_wires;
  uint64_t *o_zs_next = o_zs + 2 * nch;
  uint64_t *o_pp = o_zs_next + 2 * nch;
  uint64_t *o_quot = o_pp + 2 * (w_zs - nch);
  eval_batch(cs, 0, cs->n_cols, zeta, o_constants);
  eval_batch(&wi, 0, c->num_wires, zeta, o_wires);
  eval_batch(&zs, 0, nch, zeta, o_zs);
  eval_batch(&zs, 0, nch, zeta_next, o_zs_next);
  eval_batch(&zs, nch, w_zs - nch, zeta, o_pp);
  eval_batch(&qt, 0, w_q, zeta, o_quot);
  off += 2 * (cs->n_cols + c->num_wires + w_zs + nch + w_q);
  /* observe_openings(&openings.to_fri_openings()): the zeta batch in FRI order, then zs_next */
  challenger_observe(&ch, o_constants, 2 * (cs->n_cols + c->num_wires + nch));
  challenger_observe(&ch, o_pp, 2 * (w_zs - nch + w_q));
  challenger_observe(&ch, o_zs_next, 2 * nch);
  MARK("openings + transcript");

  /* ---- PolynomialBatch::prove_openings */
  uint64_t alpha[2];
  alpha[0] = challenger_get(&ch);
  alpha[1] = challenger_get(&ch);
  const batch_t *oracles[4] = {cs, &wi, &zs, &qt};
  const size_t m0 = cs->n_cols + c->num_wires + w_zs + w_q, m1 = nch;
  const uint64_t **polys = (const uint64_t **)malloc((m0 + m1) * sizeof(*polys));
  {
    size_t k = 0;
    for (int o = 0; o < 4; o++)
      for (size_t i = 0; i < oracles[o]->n_cols; i++) polys[k++] = oracles[o]->coeffs + i * n;
    for (size_t i = 0; i < nch; i++) polys[k++] = zs.coeffs + i * n;
  }
  uint64_t *apow = (uint64_t *)malloc(2 * (m0 + 1) * 8);
  apow[0] = 1, apow[1] = 0;
  for (size_t i = 1; i <= m0; i++) gl2_mul(apow + 2 * (i - 1), alpha, apow + 2 * i);
  uint64_t *fin = (uint64_t *)calloc(2 * n, 8), *comp = (uint64_t *)malloc(2 * n * 8), *quot = (uint64_t *)malloc(2 * n * 8);
  for (int bi = 0; bi < 2; bi++) {
    const uint64_t **pl = bi == 0 ? polys : polys + m0;
    const size_t m = bi == 0 ? m0 : m1;
    const uint64_t *pt = bi == 0 ? zeta : zeta_next;
    /* alpha.reduce_polys_base */
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; k++) {
      uint64_t a0 = 0, a1 = 0;
      for (size_t i = 0; i < m; i++) {
        const uint64_t cf = gl_canon(pl[i][k]);
        a0 = gli_add(a0, gli_mul(cf, apow[2 * i]));
        a1 = gli_add(a1, gli_mul(cf, apow[2 * i + 1]));
      }
      comp[2 * k] = a0;
      comp[2 * k + 1] = a1;
    }
    /* divide_by_linear(point) + zero pad: b_k = b_{k+1} z + c_k, quot[k-1] = b_k, quot[n-1] = 0 */
    uint64_t acc[2] = {0, 0};
    quot[2 * (n - 1)] = quot[2 * (n - 1) + 1] = 0;
    for (size_t k = n; k-- > 0;) {
      uint64_t t[2];
      gl2_mul(acc, pt, t);
      acc[0] = gli_add(t[0], comp[2 * k]);
      acc[1] = gli_add(t[1], comp[2 * k + 1]);
      if (k >= 1) quot[2 * (k - 1)] = acc[0], quot[2 * (k - 1) + 1] = acc[1];
    }
    /* alpha.shift_poly(final) by alpha^m, then final += quotient */
    const uint64_t *s = apow + 2 * m;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; k++) {
      uint64_t t[2];
      gl2_mul(fin + 2 * k, s, t);
      fin[2 * k] = gli_add(t[0], quot[2 * k]);
      fin[2 * k + 1] = gli_add(t[1], quot[2 * k + 1]);
    }
  }
  free(polys);
  free(apow);
  free(comp);
  free(quot);
  /* lde_final_poly = final_poly.lde(rate_bits); lde_final_values = coset_fft(7) */
  uint64_t *coeffs = (uint64_t *)calloc(2 * N, 8), *values = (uint64_t *)malloc(2 * N * 8);
  memcpy(coeffs, fin, 2 * n * 8);
  free(fin);
  memcpy(values, coeffs, 2 * N * 8);
  gl2_coset_fft(values, log_N, 7);
  MARK("final poly + LDE");
  /* fri_proof: commit phase */
  const size_t nl = fp->n_layers;
  uint64_t *layer_leaves[16] = {0}, *layer_digests[16] = {0};
  unsigned ab[16];
  {
    size_t ln = N;
    for (size_t l = 0; l < nl; l++) {
      ab[l] = fp->reduction_arity_bits[l];
      layer_leaves[l] = (uint64_t *)malloc(ln * 16);
      ln >>= ab[l];
      layer_digests[l] = (uint64_t *)malloc(2 * (ln - n_cap) * 32 + 32);
    }
  }
  uint64_t *caps_out = out + off;
  off += nl * 4 * n_cap;
  uint32_t sum_ab = 0;
  for (size_t l = 0; l < nl; l++) sum_ab += ab[l];
  const size_t n_final = (N >> sum_ab) >> rb;
  uint64_t *final_poly = (uint64_t *)malloc(2 * (n_final ? n_final : 1) * 8);
  fri_committed_trees(coeffs, values, N, ab, nl, rb, cap_h, &ch, caps_out, layer_leaves, layer_digests, final_poly, NULL);
  free(coeffs);
  free(values);
  MARK("fri_committed_trees");
  const uint64_t pow_witness = fri_proof_of_work(&ch, fp->proof_of_work_bits);
  MARK("proof of work");
  /* query rounds */
  for (uint32_t q = 0; q < fp->num_query_rounds; q++) {
    size_t x = (size_t)(challenger_get(&ch) & (N - 1));
    for (int o = 0; o < 4; o++) {
      const batch_t *b = oracles[o];
      memcpy(out + off, b->leaves + x * b->n_cols, b->n_cols * 8);
      off += b->n_cols;
      merkle_prove(b->digests, N, b->cap_height, x, out + off);
      off += 4 * (size_t)(log_N - b->cap_height);
    }
    size_t ln = N;
    unsigned lc = log_N;
    for (size_t l = 0; l < nl; l++) {
      ln >>= ab[l];
      lc -= ab[l];
      x >>= ab[l];
      memcpy(out + off, layer_leaves[l] + x * ((size_t)2 << ab[l]), ((size_t)2 << ab[l]) * 8);
      off += (size_t)2 << ab[l];
      merkle_prove(layer_digests[l], ln, cap_h, x, out + off);
      off += 4 * (size_t)(lc - cap_h);
    }
  }
  memcpy(out + off, final_poly, 2 * n_final * 8);
  off += 2 * n_final;
  out[off++] = pow_witness;
  for (size_t i = 0; i < n_pis; i++) out[off++] = gl_canon(public_inputs[i]);
  free(final_poly);
  for (size_t l = 0; l < nl; l++) {
    free(layer_leaves[l]);
    free(layer_digests[l]);
  }
  batch_free(&wi);
  batch_free(&zs);
  batch_free(&qt);
  MARK("query rounds");
  return off == proof_len ? off : 0;
}
