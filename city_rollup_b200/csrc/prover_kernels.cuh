// prover_kernels.cuh — the opening stage of the prover on device-resident batches.
//
// Replaces plonky2 0.2.2:
//   plonk/proof.rs   OpeningSet::new's eval_commitment (PolynomialCoeffs::to_extension().eval(z))   k_eval_polys_ext
//   fri/oracle.rs    PolynomialBatch::prove_openings: alpha.reduce_polys_base, divide_by_linear,
//                    alpha.shift_poly, final_poly.lde + coset_fft                       k_reduce_polys .. k_shift_add
//   fri/prover.rs    fri_prover_query_rounds / fri_prover_query_round (MerkleTree::get + prove per oracle and
//                    per commit-phase layer)                                            k_query_*
// Conventions pinned by the ten proofs stored in qbench_data/example.bin (tests/test_oracle_golden.py: the
// openings order, the two opening batches and the alpha bookkeeping are solved for and re-checked on all 28
// query rounds).
#pragma once
#include "fri_kernels.cuh"
#include "gl64.cuh"

namespace provk {

using gl::ext2;

// acc * z + c for a base-field coefficient c
__device__ __forceinline__ ext2 horner_step(ext2 acc, ext2 z, uint64_t c) {
  ext2 r = gl::ext_mul(acc, z);
  r.c0 = gl::add(r.c0, gl::canon(c));  // coefficients of a from_coeffs batch are the caller's raw words
  return r;
}
__device__ __forceinline__ ext2 ext_pow(ext2 b, size_t e) {
  ext2 r{1, 0};
  while (e) {
    if (e & 1) r = gl::ext_mul(r, b);
    b = gl::ext_mul(b, b);
    e >>= 1;
  }
  return r;
}

// poly(z) for a base-field polynomial of n coefficients (n a power of two) and an extension point z, by one CTA of 256
// threads; every thread returns with the value only in thread 0 (the others get garbage).  Thread t owns the `per`
// coefficients from t * per on.  Instead of a Horner chain of extension multiplications (100 instructions per
// coefficient) plus z^(t * per) by square-and-multiply in every thread (~20 more extension multiplications), the powers
// z^k, k < 16, and z^(t * per) = A[t >> 4] * B[t & 15] (A[a] = z^(16 per a), B[b] = z^(per b)) are built once per CTA by 48
// threads, and a run of 16 coefficients is two dot products c_k * Re / Im (z^k) accumulated WITHOUT reduction
// (gl::Dot160: 13 instructions per coefficient and component).  Field arithmetic is exact: same value as Horner's.
__device__ __forceinline__ ext2 eval_poly_cta(const uint64_t* __restrict__ c, size_t n, ext2 z, uint64_t* s0, uint64_t* s1,
                                              ext2* tab /* 48 entries of shared memory */) {
  const uint32_t tid = threadIdx.x;
  const size_t per = (n + 255) / 256, lo = (size_t)tid * per, hi = lo + per < n ? lo + per : n;
  if (tid < 48) {
    // group 0: z^b; group 1: (z^per)^b; group 2: (z^(16 per))^b   (per is a power of two: squarings)
    ext2 base = z;
    const uint32_t grp = tid >> 4, b = tid & 15;
    if (grp >= 1)
      for (size_t e = per; e > 1; e >>= 1) base = gl::ext_mul(base, base);
    if (grp == 2)
      for (int k = 0; k < 4; k++) base = gl::ext_mul(base, base);
    tab[tid] = ext_pow(base, b);
  }
  __syncthreads();
  ext2 acc{0, 0};
  if (lo < n) {
    const ext2 z16 = gl::ext_mul(tab[8], tab[8]);
    // chunks of 16 coefficients from the top: acc = acc * z^16 + sum_k c[s + k] z^k
    const size_t len = hi - lo, chunks = (len + 15) / 16;
    for (size_t ch = chunks; ch-- > 0;) {
      const size_t s = lo + 16 * ch, e = s + 16 < hi ? s + 16 : hi;
      gl::Dot160 d0, d1;
      for (size_t k = s; k < e; k++) {
        const uint64_t ck = gl::canon(c[k]);  // coefficients of a from_coeffs batch are the caller's raw words
        d0.add_product(ck, tab[k - s].c0);
        d1.add_product(ck, tab[k - s].c1);
      }
      if (ch + 1 < chunks) acc = gl::ext_mul(acc, z16);
      acc = gl::ext_add(acc, ext2{d0.value(), d1.value()});
    }
    acc = gl::ext_mul(acc, gl::ext_mul(tab[32 + (tid >> 4)], tab[16 + (tid & 15)]));  // * z^(tid * per)
  }
  s0[tid] = acc.c0;
  s1[tid] = acc.c1;
  __syncthreads();
  for (uint32_t off = 128; off > 0; off >>= 1) {
    if (tid < off) {
      s0[tid] = gl::add(s0[tid], s0[tid + off]);
      s1[tid] = gl::add(s1[tid], s1[tid + off]);
    }
    __syncthreads();
  }
  return ext2{s0[0], s1[0]};
}

// out[p] = poly_p(point): one CTA per polynomial (column-major coefficients, n each)
__global__ void __launch_bounds__(256) k_eval_polys_ext(const uint64_t* __restrict__ coeffs, size_t n,
                                                         const uint64_t* __restrict__ zp, uint64_t* __restrict__ out) {
  __shared__ uint64_t s0[256], s1[256];
  __shared__ ext2 tab[48];
  const ext2 z{gl::canon(zp[0]), gl::canon(zp[1])};
  const ext2 r = eval_poly_cta(coeffs + (size_t)blockIdx.x * n, n, z, s0, s1, tab);
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = r.c0;
    out[2 * blockIdx.x + 1] = r.c1;
  }
}

// pw[i] = alpha^i, i <= m (alpha read from device memory: the transcript never leaves the GPU)
__global__ void k_ext_powers(const uint64_t* __restrict__ alpha, uint32_t m, uint64_t* __restrict__ pw) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ext2 a{alpha[0], alpha[1]}, p{1, 0};
  for (uint32_t i = 0; i <= m; i++) {
    pw[2 * i] = p.c0;
    pw[2 * i + 1] = p.c1;
    p = gl::ext_mul(p, a);
  }
}

// ReducingFactor::reduce_polys_base: out[k] = sum_i alpha^i polys[i][k]   (out as two planes of n)
__global__ void __launch_bounds__(256) k_reduce_polys(const uint64_t* const* __restrict__ polys, uint32_t m, size_t n,
                                                       const uint64_t* __restrict__ pw, uint64_t* __restrict__ o0,
                                                       uint64_t* __restrict__ o1) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t a0 = 0, a1 = 0;
  for (uint32_t i = 0; i < m; i++) {
    const uint64_t c = gl::canon(polys[i][k]);
    a0 = gl::add(a0, gl::mul(c, pw[2 * i]));
    a1 = gl::add(a1, gl::mul(c, pw[2 * i + 1]));
  }
  o0[k] = a0;
  o1[k] = a1;
}

// PolynomialCoeffs::divide_by_linear(z) followed by the zero pad: q[k-1] = b_k, b_k = b_{k+1} z + c_k, q[n-1] = 0.
// One CTA of 1024 threads: local recurrences, a log-step scan of the 1024 chunk carries, local recurrences again.
__global__ void __launch_bounds__(1024) k_divide_by_linear(const uint64_t* __restrict__ c0, const uint64_t* __restrict__ c1,
                                                            size_t n, const uint64_t* __restrict__ zptr, uint64_t* __restrict__ q0,
                                                            uint64_t* __restrict__ q1) {
  __shared__ uint64_t L0[1024], L1[1024], Z0[1024], Z1[1024];
  const uint32_t tid = threadIdx.x;
  const size_t per = (n + 1023) / 1024, lo = (size_t)tid * per, hi = lo + per < n ? lo + per : n;
  const ext2 z{gl::canon(zptr[0]), gl::canon(zptr[1])};
  ext2 acc{0, 0}, zp{1, 0};
  if (lo < n)
    for (size_t k = hi; k-- > lo;) {
      acc = gl::ext_add(gl::ext_mul(acc, z), ext2{c0[k], c1[k]});
      zp = gl::ext_mul(zp, z);
    }
  L0[tid] = acc.c0, L1[tid] = acc.c1, Z0[tid] = zp.c0, Z1[tid] = zp.c1;
  __syncthreads();
  // suffix scan of the affine maps b_in -> L_t + Z_t b_in (composition towards lower t), Hillis-Steele:
  // after the scan (L_t, Z_t) maps the carry into chunk t+k-1 ... here all the way from b_n = 0, so the value
  // entering chunk t is the L of chunk t + 1.
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    ext2 l{L0[tid], L1[tid]}, p{Z0[tid], Z1[tid]};
    const bool has = tid + off < 1024;
    ext2 l2{0, 0}, p2{1, 0};
    if (has) l2 = ext2{L0[tid + off], L1[tid + off]}, p2 = ext2{Z0[tid + off], Z1[tid + off]};
    __syncthreads();
    if (has) {
      l = gl::ext_add(l, gl::ext_mul(p, l2));  // apply this chunk after the ones to its right
      p = gl::ext_mul(p, p2);
      L0[tid] = l.c0, L1[tid] = l.c1, Z0[tid] = p.c0, Z1[tid] = p.c1;
    }
    __syncthreads();
  }
  const ext2 carry_in = tid + 1 < 1024 ? ext2{L0[tid + 1], L1[tid + 1]} : ext2{0, 0};
  __syncthreads();
  if (lo >= n) return;
  acc = carry_in;
  for (size_t k = hi; k-- > lo;) {
    acc = gl::ext_add(gl::ext_mul(acc, z), ext2{c0[k], c1[k]});
    if (k >= 1) q0[k - 1] = acc.c0, q1[k - 1] = acc.c1;
  }
  if (hi == n) q0[n - 1] = 0, q1[n - 1] = 0;
}

// alpha.shift_poly(final) ; final += quotient:   f = f * s + q  with s = pw[2 m], pw[2 m + 1]
__global__ void __launch_bounds__(256) k_shift_add(uint64_t* __restrict__ f0, uint64_t* __restrict__ f1,
                                                    const uint64_t* __restrict__ q0, const uint64_t* __restrict__ q1, size_t n,
                                                    const uint64_t* __restrict__ s) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  ext2 r = gl::ext_mul(ext2{f0[k], f1[k]}, ext2{s[0], s[1]});
  f0[k] = gl::add(r.c0, q0[k]);
  f1[k] = gl::add(r.c1, q1[k]);
}

// out = g * z for a base-field g (zeta_next = g * zeta)
__global__ void k_ext_scale(const uint64_t* __restrict__ z, uint64_t g, uint64_t* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out[0] = gl::mul(z[0], g);
  out[1] = gl::mul(z[1], g);
}

// ---------------------------------------------------------------------------------------------- query rounds
// x_index of query q = challenge[q] mod lde_size, shifted right by `shift` (the commit-phase layers)
__device__ __forceinline__ size_t query_index(const uint64_t* chal, uint32_t q, uint32_t log_lde, uint32_t shift) {
  return (size_t)((gl::canon(chal[q]) & (((uint64_t)1 << log_lde) - 1)) >> shift);
}

// MerkleTree::get for a column-major batch: out[q * stride_out + c] = data[c * col_stride + x_q]
__global__ void k_query_leaf_colmajor(const uint64_t* __restrict__ data, size_t col_stride, uint32_t n_cols,
                                      const uint64_t* __restrict__ chal, uint32_t log_lde, uint64_t* __restrict__ out,
                                      size_t stride_out) {
  const uint32_t q = blockIdx.x;
  const size_t x = query_index(chal, q, log_lde, 0);
  for (uint32_t c = threadIdx.x; c < n_cols; c += blockDim.x) out[(size_t)q * stride_out + c] = data[(size_t)c * col_stride + x];
}
// MerkleTree::get for row-major leaves (FRI layers)
__global__ void k_query_leaf_rowmajor(const uint64_t* __restrict__ leaves, uint32_t leaf_len, const uint64_t* __restrict__ chal,
                                      uint32_t log_lde, uint32_t shift, uint64_t* __restrict__ out, size_t stride_out) {
  const uint32_t q = blockIdx.x;
  const size_t x = query_index(chal, q, log_lde, shift);
  for (uint32_t c = threadIdx.x; c < leaf_len; c += blockDim.x) out[(size_t)q * stride_out + c] = leaves[x * leaf_len + c];
}
// MerkleTree::prove: siblings of leaf x_q for layers 0 .. L-1 (levels: leaf level first)
__global__ void k_query_siblings(const uint64_t* __restrict__ levels, size_t n_leaves, uint32_t L, const uint64_t* __restrict__ chal,
                                 uint32_t log_lde, uint32_t shift, uint64_t* __restrict__ out, size_t stride_out) {
  const uint32_t q = blockIdx.x;
  const size_t x = query_index(chal, q, log_lde, shift);
  for (uint32_t t = threadIdx.x; t < 4 * L; t += blockDim.x) {
    const uint32_t i = t >> 2, w = t & 3;
    const size_t off = 2 * n_leaves - 2 * (n_leaves >> i);
    out[(size_t)q * stride_out + t] = levels[4 * (off + ((x >> i) ^ 1)) + w];
  }
}

}  // namespace provk
