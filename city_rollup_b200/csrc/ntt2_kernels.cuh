// ntt2_kernels.cuh — register-resident radix-16 Goldilocks NTT for 2^12 <= n <= 2^20 (the sizes the City
// Rollup circuits and the BASELINE.json shapes use); smaller / larger transforms stay on ntt_kernels.cuh.
//
// Same contract as ntt_kernels.cuh (plonky2_field 0.2.2 fft.rs / polynomial/mod.rs):
// forward X[r] = sum_k x[k] w_n^(k r), w_n = G^(2^(32 - log n)); decimation in frequency, so a transform
// leaves its result in bit-reversed order — the order plonky2's Merkle leaves want.
//
// Design (B200).  The butterflies are INT32-pipe bound (about 300 integer instructions per element against a
// machine balance of 43 per 8 bytes), so the kernels are built to keep every element in
// registers for 4 layers at a time and to touch shared / global memory once per 4 layers:
//   * Dif<4>: a 16-point DIF on 16 registers.  Its internal twiddles are powers of w_16 = 2^156 = -2^60
//     (2 has order 192 in Goldilocks, w_64 = 2^39), i.e. compile-time constants with one or two set bits;
//     negative ones are folded into the subtraction order.
//   * k_row4096: one CTA = one contiguous row of 4096 elements = three Dif<4> stages; stage 1 reads global
//     memory straight into registers (coalesced, stride 256), stages exchange through a padded 34 KB tile
//     (conflict-free for 64-bit accesses), results leave through the tile so that the global stores are
//     coalesced.  Twiddles between stages come from two small tables laid out so that a warp reads them
//     contiguously.  For n = 4096 the LDE loops over the 2^rate_bits cosets inside the CTA: the
//     coefficients are read once and prescaled from a per-coset table.
//   * k_strided<D>: for n > 4096, one radix-2^D step (D <= 4) over elements m/2^D apart, twiddled by
//     w_m^(r0 k0) from a table laid out [r0][k0] (coalesced), output slot = bit-reversed r0.  The first step
//     of an LDE also loops over the cosets, so the coefficient matrix is read from HBM once, not 2^rate
//     times.
// All values are kept canonical (< p): add 7, sub 5, mul 18 + 4 integer instructions.
#pragma once
#include "gl64.cuh"

namespace ntt2 {

__device__ __forceinline__ uint64_t mulc(uint64_t a, uint64_t b) { return gl::canon(gl::mul_nc(a, b)); }

// a + b mod p for canonical a, b, canonical result: a - (p - b), + p on borrow
__device__ __forceinline__ uint64_t addc(uint64_t a, uint64_t b) {
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u32 nl,nh,m;\n\t"
      "sub.cc.u32 nl, 1, %4;\n\t"
      "subc.u32 nh, 0xFFFFFFFF, %5;\n\t"
      "sub.cc.u32 %0, %2, nl;\n\t"
      "subc.cc.u32 %1, %3, nh;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
  return gl::pack(r0, r1);
}
// a - b mod p for canonical a, b, canonical result
__device__ __forceinline__ uint64_t subc(uint64_t a, uint64_t b) { return gl::sub_nc(a, b); }

// 2^e mod p for e < 96 (2^64 = 2^32 - 1)
__host__ __device__ constexpr uint64_t pow2_mod_p(int e) {
  return e < 64 ? (1ull << e) : ((1ull << (e - 32)) - (1ull << (e - 64)));
}
__host__ __device__ constexpr uint32_t brev_c(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

// x * 2^E mod p for canonical x and 0 < E < 96 (E is a compile-time constant after unrolling), canonical result.
// x << (E mod 32) is three 32-bit words y0, y1, y2 (funnel shifts) that land on word positions q .. q+2,
// q = E / 32, and 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 reduce them with one or two modular add/subs:
// ~15 instructions instead of the ~34 a general multiplication by the constant costs.
__device__ __forceinline__ uint64_t mul_pow2c(uint64_t x, int E) {
  const int r = E & 31, q = E >> 5;
  const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
  const uint32_t y0 = xl << r;
  const uint32_t y1 = r ? __funnelshift_l(xl, xh, r) : xh;
  const uint32_t y2 = r ? (xh >> (32 - r)) : 0u;
  auto eps_times = [](uint32_t v) { return ((uint64_t)v << 32) - v; };  // v * (2^32 - 1), canonical
  uint64_t res;
  if (q == 0)
    res = gl::add_nc(gl::pack(y0, y1), eps_times(y2));                                   // y0 + 2^32 y1 + 2^64 y2
  else if (q == 1)
    res = gl::sub_nc(gl::add_nc(gl::pack(0u, y0), eps_times(y1)), (uint64_t)y2);           // 2^32 y0 + 2^64 y1 + 2^96 y2
  else
    res = gl::sub_nc(eps_times(y0), gl::pack(y1, y2));                                     // 2^64 y0 + 2^96 y1 + 2^128 y2
  return gl::canon(res);
}

// In-register DIF of size 2^LOGR (LOGR <= 6): slot `pos` ends up holding output index brev(pos).
// w_{2^k} = 2^(39 * 2^(6-k)) mod p, so w_{2h}^j = +-2^E with E known at compile time.
template <int LOGR>
struct Dif {
  static constexpr int R = 1 << LOGR;
  __device__ __forceinline__ static void run(uint64_t (&x)[R]) {
#pragma unroll
    for (int l = 0; l < LOGR; l++) {
      const int half = R >> (l + 1);
#pragma unroll
      for (int blk = 0; blk < (1 << l); blk++) {
#pragma unroll
        for (int j = 0; j < half; j++) {
          const int i0 = blk * 2 * half + j, i1 = i0 + half;
          const int E = (39 * (64 / (2 * half)) * j) % 192;
          const uint64_t a = x[i0], b = x[i1];
          x[i0] = addc(a, b);
          if (E == 0)
            x[i1] = subc(a, b);
          else if (E < 96)
            x[i1] = mul_pow2c(subc(a, b), E);
          else
            x[i1] = mul_pow2c(subc(b, a), E - 96);
        }
      }
    }
  }
};

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// ---------------------------------------------------------------------------------------------- tables
struct RootTables {
  const uint64_t* r_lo;  // G^j, j < 65536
  const uint64_t* r_hi;  // G^(65536 i)
};
__device__ __forceinline__ uint64_t root_pow(const RootTables& t, uint32_t log_m, uint64_t e) {
  uint32_t E = (uint32_t)(e << (32 - log_m));
  return mulc(t.r_hi[E >> 16], t.r_lo[E & 0xFFFFu]);
}
// out[r0 * (m >> d) + k0] = w_m^(r0 * k0), r0 < 2^d, k0 < m >> d
__global__ void k_build_tw(RootTables t, uint32_t log_m, uint32_t d, uint64_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >> log_m) return;
  uint32_t lmp = log_m - d;
  uint64_t r0 = i >> lmp, k0 = i & (((size_t)1 << lmp) - 1);
  out[i] = root_pow(t, log_m, (r0 * k0) & (((uint64_t)1 << log_m) - 1));
}

// ---------------------------------------------------------------------------------------------- strided step
struct StridedParams {
  const uint64_t* in;
  size_t in_col_stride;
  uint64_t* out;
  size_t out_col_stride;
  size_t out_coset_stride;
  const uint64_t* tw;   // [2^D][m >> D]
  const uint64_t* cp;   // PRESCALE: [cosets][n] powers of the coset shifts
  uint32_t log_m;       // size of the sub-transforms this step splits
  uint32_t log_n;       // PRESCALE: transform size (= log_m), stride of cp
  uint32_t log_cosets;  // PRESCALE: cosets looped over inside the thread
};

// grid = (items / 256, n_cols); one thread = one (sub-transform, k0) pair = 2^D elements m >> D apart.
// Without PRESCALE the step runs in place over `total` contiguous elements of each column.
template <int D, bool PRESCALE>
__global__ void __launch_bounds__(256) k_strided(StridedParams P) {
  constexpr int R = 1 << D;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  const uint32_t lmp = P.log_m - D;
  const size_t mp = (size_t)1 << lmp;
  const size_t k0 = idx & (mp - 1);
  const size_t base = ((idx >> lmp) << P.log_m) + k0;
  const uint64_t* in = P.in + (size_t)blockIdx.y * P.in_col_stride + base;
  uint64_t xin[R];
#pragma unroll
  for (int k1 = 0; k1 < R; k1++) xin[k1] = gl::canon(in[(size_t)k1 << lmp]);
  uint64_t w[R];
#pragma unroll
  for (int pos = 1; pos < R; pos++) w[pos] = P.tw[((size_t)brev_c(pos, D) << lmp) + k0];
  const uint32_t n_t = PRESCALE ? (1u << P.log_cosets) : 1u;
#pragma unroll 1
  for (uint32_t t = 0; t < n_t; t++) {
    uint64_t x[R];
    if (PRESCALE) {
      const uint64_t* cp = P.cp + ((size_t)t << P.log_n) + base;
#pragma unroll
      for (int k1 = 0; k1 < R; k1++) x[k1] = mulc(xin[k1], cp[(size_t)k1 << lmp]);
    } else {
#pragma unroll
      for (int k1 = 0; k1 < R; k1++) x[k1] = xin[k1];
    }
    Dif<D>::run(x);
    uint64_t* out = P.out + (size_t)blockIdx.y * P.out_col_stride + base;
    if (PRESCALE) out += (size_t)brev(t, P.log_cosets) * P.out_coset_stride;
    out[0] = x[0];
#pragma unroll
    for (int pos = 1; pos < R; pos++) out[(size_t)pos << lmp] = mulc(x[pos], w[pos]);
  }
}

// ---------------------------------------------------------------------------------------------- 4096-point rows
struct RowParams {
  const uint64_t* in;
  size_t in_col_stride;
  uint64_t* out;
  size_t out_col_stride;
  size_t out_coset_stride;
  const uint64_t* t1;   // [16][256]: w_4096^(a q)
  const uint64_t* t2;   // [16][16]:  w_256^(b c)
  const uint64_t* cp;   // PRESCALE: [cosets][4096]
  uint32_t log_cosets;  // PRESCALE: cosets looped over inside the CTA
  uint32_t log_R;       // natural modes: the column holds 2^log_R rows (n = 4096 << log_R)
  uint64_t scale;       // MODE 2: 1/n
};

__device__ __forceinline__ uint32_t pad16(uint32_t i) { return i + (i >> 4); }

// MODE 0: row in, row out in bit-reversed (DIF) order — LDE / in-place last step.
// MODE 1: natural order: the row is sub-transform brev(row) of a size-n DIF; X[r_low + (r_high << log_R)].
// MODE 2: MODE 1 + index reversal (n - r) mod n and scaling: plonky2's ifft.
// grid = (rows per column, n_cols); 256 threads, 16 elements per thread.
template <int MODE, bool PRESCALE>
#ifndef P2B_ROW_MINB
#define P2B_ROW_MINB 3
#endif
__global__ void __launch_bounds__(256, P2B_ROW_MINB) k_row4096(RowParams P) {
  __shared__ uint64_t sm[4096 + 256];
  const uint32_t tid = threadIdx.x;
  const uint32_t blk = tid >> 4, c = tid & 15;
  const size_t row = blockIdx.x;
  const uint64_t* in = P.in + (size_t)blockIdx.y * P.in_col_stride + row * 4096;
  uint64_t xin[16];
#pragma unroll
  for (int a = 0; a < 16; a++) xin[a] = gl::canon(in[a * 256 + tid]);
  const uint32_t n_t = PRESCALE ? (1u << P.log_cosets) : 1u;
#pragma unroll 1
  for (uint32_t t = 0; t < n_t; t++) {
    uint64_t x[16];
    if (PRESCALE) {
      const uint64_t* cp = P.cp + (size_t)t * 4096 + tid;
#pragma unroll
      for (int a = 0; a < 16; a++) x[a] = mulc(xin[a], cp[a * 256]);
    } else {
#pragma unroll
      for (int a = 0; a < 16; a++) x[a] = xin[a];
    }
    // stage 1: over a (stride 256); slot pos -> block pos of the tile, twiddle w_4096^(brev4(pos) * tid)
    Dif<4>::run(x);
    if (t > 0) __syncthreads();  // the previous coset's copy-out has finished
    sm[pad16(tid)] = x[0];
#pragma unroll
    for (int pos = 1; pos < 16; pos++) sm[pad16(pos * 256 + tid)] = mulc(x[pos], P.t1[brev_c(pos, 4) * 256 + tid]);
    __syncthreads();
    // stage 2: inside block blk, over b (stride 16), lane c; twiddle w_256^(brev4(pos) * c)
#pragma unroll
    for (int b = 0; b < 16; b++) x[b] = sm[pad16(blk * 256 + b * 16 + c)];
    Dif<4>::run(x);
    sm[pad16(blk * 256 + c)] = x[0];
#pragma unroll
    for (int pos = 1; pos < 16; pos++) sm[pad16(blk * 256 + pos * 16 + c)] = mulc(x[pos], P.t2[brev_c(pos, 4) * 16 + c]);
    __syncthreads();
    // stage 3: 16 contiguous elements per thread
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = sm[tid * 17 + i];
    Dif<4>::run(x);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; i++) sm[tid * 17 + i] = x[i];
      __syncthreads();
      uint64_t* out = P.out + (size_t)blockIdx.y * P.out_col_stride + row * 4096;
      if (PRESCALE) out += (size_t)brev(t, P.log_cosets) * P.out_coset_stride;
#pragma unroll
      for (int i = 0; i < 16; i++) out[i * 256 + tid] = sm[pad16(i * 256 + tid)];
    } else {
      // slot (blk, pos2 = c, pos3) holds r_high = brev4(blk) + 16 brev4(c) + 256 brev4(pos3)
      const uint32_t rh0 = brev(blk, 4) + 16 * brev(c, 4);
      __syncthreads();  // every thread has read its stage-3 inputs
#pragma unroll
      for (int i = 0; i < 16; i++) {
        uint64_t v = x[i];
        if (MODE == 2) v = mulc(v, P.scale);
        sm[pad16(rh0 + 256 * brev_c(i, 4))] = v;
      }
      __syncthreads();
      const size_t n = (size_t)4096 << P.log_R;
      const size_t r_low = brev((uint32_t)row, P.log_R);
      uint64_t* out = P.out + (size_t)blockIdx.y * P.out_col_stride;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t rh = i * 256 + tid;
        size_t idx = r_low + ((size_t)rh << P.log_R);
        if (MODE == 2) idx = (n - idx) & (n - 1);
        out[idx] = sm[pad16(rh)];
      }
    }
  }
}

// t1[a * 256 + q] = w_4096^(a q); t2[b * 16 + c] = w_256^(b c)
__global__ void k_build_row_tables(RootTables t, uint64_t* __restrict__ t1, uint64_t* __restrict__ t2) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4096) t1[i] = root_pow(t, 12, (uint64_t)(i >> 8) * (i & 255));
  if (i < 256) t2[i] = root_pow(t, 8, (uint64_t)(i >> 4) * (i & 15));
}

}  // namespace ntt2
