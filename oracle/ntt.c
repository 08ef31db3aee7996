/* ntt.c — FFT / coset FFT / LDE and PolynomialBatch::{from_values,from_coeffs}.
 * TEST INFRASTRUCTURE (see p2oracle.h).
 * Restates plonky2_field 0.2.2 fft.rs (fft = bit-reverse then radix-2 DIT, natural-order output;
 * ifft = fft, scale 1/n, reverse indices 1..n), polynomial/mod.rs (lde = zero pad; coset_fft =
 * scale coeff i by shift^i then fft) and plonky2 0.2.2 fri/oracle.rs (from_values -> ifft ->
 * from_coeffs -> lde_values -> transpose -> reverse_index_bits_in_place -> MerkleTree::new).
 * None of these is on disk.  The worker circuits reach them through
 * circuit_data.prove(pw), e.g. city_common_circuit/src/proof_minifier/pm_core.rs:151.
 * PARITY UNPINNED by reference fixtures (no test in the reference asserts an LDE value or a cap);
 * exact field arithmetic makes any correct NTT produce the same canonical values. */
#include <stdlib.h>
#include <string.h>

#include "gl_inline.h"

/* in-place natural->natural forward transform: X[r] = sum_k x[k] w^(k r) */
static void fft_with_root(uint64_t *a, unsigned log_n, uint64_t w) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) {
      uint64_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  for (size_t i = 0; i < n; i++) a[i] = gl_canon(a[i]);
  if (log_n == 0) return;
  uint64_t *tw = (uint64_t *)malloc((n / 2) * sizeof(uint64_t));
  tw[0] = 1;
  for (size_t i = 1; i < n / 2; i++) tw[i] = gli_mul(tw[i - 1], w);
  for (unsigned s = 1; s <= log_n; s++) {
    size_t m = (size_t)1 << s, half = m >> 1, step = n / m;
    for (size_t k = 0; k < n; k += m)
      for (size_t j = 0; j < half; j++) {
        uint64_t u = a[k + j], t = gli_mul(a[k + j + half], tw[j * step]);
        a[k + j] = gli_add(u, t);
        a[k + j + half] = gli_sub(u, t);
      }
  }
  free(tw);
}

void gl_fft(uint64_t *a, unsigned log_n) { fft_with_root(a, log_n, gl_root_of_unity(log_n)); }

void gl_ifft(uint64_t *a, unsigned log_n) {
  size_t n = (size_t)1 << log_n;
  fft_with_root(a, log_n, gl_inv(gl_root_of_unity(log_n)));
  uint64_t ninv = gl_inv((uint64_t)n);
  for (size_t i = 0; i < n; i++) a[i] = gli_mul(a[i], ninv);
}

void gl_coset_fft(uint64_t *a, unsigned log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  uint64_t s = 1;
  for (size_t i = 0; i < n; i++) {
    a[i] = gli_mul(gl_canon(a[i]), s);
    s = gli_mul(s, shift);
  }
  gl_fft(a, log_n);
}

void gl_coset_ifft(uint64_t *a, unsigned log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  gl_ifft(a, log_n);
  uint64_t si = gl_inv(shift), s = 1;
  for (size_t i = 0; i < n; i++) {
    a[i] = gli_mul(a[i], s);
    s = gli_mul(s, si);
  }
}

/* extension elements are interleaved (c0,c1); the shift and the roots are base-field, so the
 * transform acts on the two components independently */
void gl2_coset_fft(uint64_t *a, unsigned log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  uint64_t *t = (uint64_t *)malloc(n * sizeof(uint64_t));
  for (int comp = 0; comp < 2; comp++) {
    for (size_t i = 0; i < n; i++) t[i] = a[2 * i + comp];
    gl_coset_fft(t, log_n, shift);
    for (size_t i = 0; i < n; i++) a[2 * i + comp] = t[i];
  }
  free(t);
}

void batch_from_coeffs(const uint64_t *const *cols, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       unsigned cap_height, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out) {
  size_t n = (size_t)1 << log_n, N = n << rate_bits;
  unsigned log_N = log_n + rate_bits;
  uint64_t *leaves = leaves_out ? leaves_out : (uint64_t *)malloc(N * n_cols * sizeof(uint64_t));
  uint32_t *rev = (uint32_t *)malloc(N * sizeof(uint32_t)); /* reverse_index_bits permutation, shared by all columns */
  for (size_t j = 0; j < N; j++) rev[j] = (uint32_t)bitrev(j, log_N);
#pragma omp parallel
  {
    uint64_t *buf = (uint64_t *)malloc(N * sizeof(uint64_t));
#pragma omp for schedule(dynamic)
    for (size_t c = 0; c < n_cols; c++) {
      memcpy(buf, cols[c], n * sizeof(uint64_t));
      memset(buf + n, 0, (N - n) * sizeof(uint64_t)); /* lde(rate_bits) */
      gl_coset_fft(buf, log_N, 7);                     /* coset_fft(F::coset_shift()) */
      for (size_t j = 0; j < N; j++) leaves[j * n_cols + c] = buf[rev[j]];
    }
    free(buf);
  }
  free(rev);
  if (digests_out || cap_out) {
    size_t n_cap = (size_t)1 << cap_height;
    uint64_t *dg = digests_out ? digests_out : (uint64_t *)malloc(2 * (N - n_cap) * 32 + 32);
    uint64_t capbuf[4 * 256];
    uint64_t *cp = cap_out ? cap_out : capbuf;
    merkle_tree_new(leaves, N, n_cols, cap_height, dg, cp);
    if (!digests_out) free(dg);
  }
  if (!leaves_out) free(leaves);
}

void batch_from_values(const uint64_t *const *cols, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       unsigned cap_height, uint64_t *coeffs_out, uint64_t *leaves_out,
                       uint64_t *digests_out, uint64_t *cap_out) {
  size_t n = (size_t)1 << log_n;
  uint64_t *coeffs = coeffs_out ? coeffs_out : (uint64_t *)malloc(n * n_cols * sizeof(uint64_t));
  const uint64_t **ptrs = (const uint64_t **)malloc(n_cols * sizeof(uint64_t *));
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < n_cols; c++) {
    memcpy(coeffs + c * n, cols[c], n * sizeof(uint64_t));
    gl_ifft(coeffs + c * n, log_n);
    ptrs[c] = coeffs + c * n;
  }
  batch_from_coeffs(ptrs, n_cols, log_n, rate_bits, cap_height, leaves_out, digests_out, cap_out);
  free(ptrs);
  if (!coeffs_out) free(coeffs);
}
