// ntt_kernels.cuh — batched Goldilocks NTT / inverse NTT / coset low-degree extension.
//
// Replaces plonky2_field 0.2.2 fft.rs (fft_dispatch / ifft), polynomial/mod.rs (lde, coset_fft) and
// the transpose + reverse_index_bits_in_place of plonky2 fri/oracle.rs::PolynomialBatch::lde_values
// (SURVEY.md A.3/A.4).  Conventions reproduced exactly: forward X[r] = sum_k x[k] w^(k r) with
// w = primitive_root_of_unity(log n) = G^(2^(32-log n)), G = 7^((p-1)/2^32); ifft(v)[k] =
// fft(v)[(n-k) mod n] / n; coset shift 7; leaf j of a batch holds LDE row bitrev(j).
//
// Design (B200): a transform of size n = R*S is two passes over HBM/L2, each a radix-2 DIF done
// entirely in shared memory on a coalesced tile:
//   pass A (k_ntt_cols): tile = R rows x T contiguous elements (row stride S); size-R DIF down the
//          rows for T independent lanes, twiddle w_n^(k0*r0), rows stored in bit-reversed or natural order;
//   pass B (k_ntt_rows): tile = TR rows of S contiguous elements; size-S DIF along each row.
// Decimation in frequency leaves results in bit-reversed order, which is exactly the order plonky2's
// Merkle leaves want, so the LDE needs no transpose and no separate bit-reversal pass: the LDE of
// rate 2^r is 2^r independent size-n coset NTTs (shift 7*w_{rn}^t), written straight into the
// column-major leaf-ordered LDE buffer (coset t lands in block bitrev_r(t)).
// n <= 2^12 is a single pass (R = 1).  Both passes support sizes up to 2^12, i.e. n <= 2^24.
#pragma once
#include "gl64.cuh"

namespace nttk {

struct Tables {
  const uint64_t* w12;   // w_4096^j, j < 2048
  const uint64_t* r_lo;  // G^j, j < 65536           (G = primitive 2^32-th root)
  const uint64_t* r_hi;  // G^(65536 i), i < 65536
};

// w_{2^log_n}^e, e < 2^log_n
__device__ __forceinline__ uint64_t root_pow(const Tables& t, uint32_t log_n, uint64_t e) {
  uint32_t E = (uint32_t)(e << (32 - log_n));
  return gl::mul(t.r_hi[E >> 16], t.r_lo[E & 0xFFFFu]);
}

// per-call power tables of the coset shifts: lo[t][j] = s_t^j (j < 4096), hi[t][i] = s_t^(4096 i)
struct CosetPow {
  const uint64_t* lo;
  const uint64_t* hi;
  uint32_t hi_count;  // entries per coset in hi
};
__device__ __forceinline__ uint64_t coset_pow(const CosetPow& c, uint32_t t, size_t k) {
  uint64_t a = c.lo[(size_t)t * 4096 + (k & 4095)];
  if (c.hi_count <= 1) return a;
  return gl::mul(a, c.hi[(size_t)t * c.hi_count + (k >> 12)]);
}

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// base^i for i < count, written to out[i]; used for every power table
__global__ void k_pow_table(uint64_t base, size_t count, uint64_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = gl::pow(gl::canon(base), i);
}

struct ColsParams {
  const uint64_t* in;
  size_t in_poly_stride;
  uint64_t* out;
  size_t out_poly_stride;
  size_t out_coset_stride;
  uint32_t log_n, lr, ls, log_T;
  uint32_t log_cosets;   // number of cosets = 1 << log_cosets (grid.z)
  uint32_t natural_rows; // 0: row pos (bit-reversed r0) ; 1: row r0
  uint32_t prescale;     // multiply input k by coset_pow(t, k)
  CosetPow cp;
  Tables tb;
};

// pass A: grid = (S / T, n_polys, n_cosets); dynamic smem = R * T * 8 bytes
__global__ void __launch_bounds__(256) k_ntt_cols(ColsParams P) {
  extern __shared__ uint64_t sm[];
  const uint32_t R = 1u << P.lr, T = 1u << P.log_T;
  const size_t S = (size_t)1 << P.ls;
  const uint32_t poly = blockIdx.y, t = blockIdx.z;
  const size_t k0b = (size_t)blockIdx.x << P.log_T;
  const uint64_t* in = P.in + (size_t)poly * P.in_poly_stride;
  const uint32_t total = R << P.log_T;
  for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
    uint32_t k1 = idx >> P.log_T, tt = idx & (T - 1);
    size_t k = (size_t)k1 * S + k0b + tt;
    uint64_t v = gl::canon(in[k]);
    if (P.prescale) v = gl::mul(v, coset_pow(P.cp, t, k));
    sm[idx] = v;
  }
  __syncthreads();
  // radix-2 DIF down the rows; lanes tt are independent
  for (uint32_t layer = 0; layer < P.lr; layer++) {
    const uint32_t lh = P.lr - 1 - layer;  // log2(half)
    const uint32_t half = 1u << lh;
    for (uint32_t idx = threadIdx.x; idx < (total >> 1); idx += blockDim.x) {
      uint32_t tt = idx & (T - 1), b = idx >> P.log_T;
      uint32_t j = b & (half - 1);
      uint32_t i0 = ((b >> lh) << (lh + 1)) | j, i1 = i0 + half;
      uint64_t a = sm[(i0 << P.log_T) + tt], c = sm[(i1 << P.log_T) + tt];
      uint64_t w = P.tb.w12[j << (11 - lh)];  // w_{2*half}^j
      sm[(i0 << P.log_T) + tt] = gl::add(a, c);
      sm[(i1 << P.log_T) + tt] = gl::mul(gl::sub(a, c), w);
    }
    __syncthreads();
  }
  uint64_t* out = P.out + (size_t)poly * P.out_poly_stride + (size_t)brev(t, P.log_cosets) * P.out_coset_stride;
  for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
    uint32_t pos = idx >> P.log_T, tt = idx & (T - 1);
    uint32_t r0 = brev(pos, P.lr);
    size_t k0 = k0b + tt;
    uint64_t v = gl::mul(sm[idx], root_pow(P.tb, P.log_n, (uint64_t)k0 * r0));
    size_t row = P.natural_rows ? r0 : pos;
    out[row * S + k0] = v;
  }
}

struct RowsParams {
  const uint64_t* in;
  size_t in_poly_stride;
  uint64_t* out;
  size_t out_poly_stride;
  size_t out_coset_stride;
  uint32_t log_n;       // full transform size (natural-order addressing)
  uint32_t lr, ls;      // n = 2^lr rows x 2^ls
  uint32_t log_TR;      // rows per CTA
  uint32_t log_cosets;
  uint32_t mode;        // 0: in-place order (row, pos) ; 1: natural order ; 2: natural + inverse (index n-k, scale)
  uint32_t prescale;    // only with lr == 0
  uint64_t scale;       // 1/n for mode 2
  CosetPow cp;
  Tables tb;
};

// pass B / single pass: grid = (rows_per_poly / TR, n_polys, n_cosets); smem = TR * (S + pad) * 8
__global__ void __launch_bounds__(256) k_ntt_rows(RowsParams P) {
  extern __shared__ uint64_t sm[];
  const uint32_t S = 1u << P.ls, TR = 1u << P.log_TR;
  const uint32_t SP = TR > 1 ? S + 1 : S;  // padded row stride: TR rows hit different banks
  const uint32_t poly = blockIdx.y, t = blockIdx.z;
  const size_t row0 = (size_t)blockIdx.x << P.log_TR;
  const uint64_t* in = P.in + (size_t)poly * P.in_poly_stride + row0 * S;
  const uint32_t total = TR << P.ls;
  for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
    uint32_t rl = idx >> P.ls, q = idx & (S - 1);
    uint64_t v = gl::canon(in[idx]);
    if (P.prescale) v = gl::mul(v, coset_pow(P.cp, t, q));
    sm[rl * SP + q] = v;
  }
  __syncthreads();
  for (uint32_t layer = 0; layer < P.ls; layer++) {
    const uint32_t lh = P.ls - 1 - layer;
    const uint32_t half = 1u << lh;
    for (uint32_t idx = threadIdx.x; idx < (total >> 1); idx += blockDim.x) {
      uint32_t rl = idx >> (P.ls - 1), b = idx & ((S >> 1) - 1);
      uint32_t j = b & (half - 1);
      uint32_t i0 = rl * SP + (((b >> lh) << (lh + 1)) | j), i1 = i0 + half;
      uint64_t a = sm[i0], c = sm[i1];
      uint64_t w = P.tb.w12[j << (11 - lh)];
      sm[i0] = gl::add(a, c);
      sm[i1] = gl::mul(gl::sub(a, c), w);
    }
    __syncthreads();
  }
  uint64_t* out = P.out + (size_t)poly * P.out_poly_stride + (size_t)brev(t, P.log_cosets) * P.out_coset_stride;
  if (P.mode == 0) {
    for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
      uint32_t rl = idx >> P.ls, q = idx & (S - 1);
      out[row0 * S + idx] = sm[rl * SP + q];
    }
  } else {
    // natural order: X index = r0 + R * r1 with r1 = bitrev(pos); consecutive threads take consecutive r0
    const size_t n = (size_t)1 << P.log_n;
    for (uint32_t o = threadIdx.x; o < total; o += blockDim.x) {
      uint32_t rl = o & (TR - 1), r1 = o >> P.log_TR;
      uint32_t pos = brev(r1, P.ls);
      uint64_t v = sm[rl * SP + pos];
      size_t idx = (row0 + rl) + ((size_t)r1 << P.lr);
      if (P.mode == 2) {
        idx = (n - idx) & (n - 1);
        v = gl::mul(v, P.scale);
      }
      out[idx] = v;
    }
  }
}

}  // namespace nttk
