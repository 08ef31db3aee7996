// fri_kernels.cuh — device-resident Challenger, FRI arity folding, proof-of-work grinding.
//
// Replaces plonky2 0.2.2 iop/challenger.rs (Challenger<F, PoseidonHash>), fri/prover.rs
// (fri_committed_trees' fold `reduce_with_powers(chunk, beta)`, fri_proof_of_work) — SURVEY.md A.7/A.9;
// FRI parameters as dumped at city_common_circuit/src/circuits/zk_signature2/mod.rs:38-50.
// The transcript state lives in HBM and is advanced by one-warp kernels (a warp-cooperative permutation), so the commit loop
// (tree -> observe cap -> beta -> fold -> coset NTT) never synchronises with the host.
#pragma once
#include "poseidon.cuh"

namespace frik {

// Challenger state: sponge_state[12] | n_in | input_buffer[8] | n_out | output_buffer[8]
constexpr int CH_WORDS = 30;
constexpr int CH_NIN = 12, CH_IN = 13, CH_NOUT = 21, CH_OUT = 22;

// Challenger::duplexing on a state held in shared memory, one permutation shared by the warp
// (poseidon::coop_permute_nc); all 32 lanes call
__device__ __forceinline__ void duplexing_coop(uint64_t* st, uint32_t lane) {
  const uint32_t n_in = (uint32_t)st[CH_NIN];
  uint64_t s = lane < 12 ? st[lane] : 0;
  if (lane < n_in) s = st[CH_IN + lane];  // overwrite mode (n_in <= 8)
  __syncwarp();
  s = gl::canon(poseidon::coop_permute_nc(s, lane));
  if (lane < 12) st[lane] = s;
  if (lane < 8) st[CH_OUT + lane] = s;
  if (lane == 0) {
    st[CH_NIN] = 0;
    st[CH_NOUT] = 8;
  }
  __syncwarp();
}

// Challenger::observe_elements; launch <<<1, 32>>>
__global__ void k_challenger_observe(uint64_t* __restrict__ g_st, const uint64_t* __restrict__ elems, size_t n) {
  __shared__ uint64_t st[32];
  const uint32_t lane = threadIdx.x;
  if (lane < CH_WORDS) st[lane] = g_st[lane];
  __syncwarp();
  size_t i = 0;
  while (i < n) {
    const uint32_t k = (uint32_t)st[CH_NIN];
    const uint32_t take = (uint32_t)((n - i) < (size_t)(8 - k) ? (n - i) : (size_t)(8 - k));
    __syncwarp();
    if (lane < take) st[CH_IN + k + lane] = gl::canon(elems[i + lane]);
    if (lane == 0) {
      st[CH_NOUT] = 0;
      st[CH_NIN] = k + take;
    }
    __syncwarp();
    i += take;
    if (k + take == 8) duplexing_coop(st, lane);
  }
  __syncwarp();
  if (lane < CH_WORDS) g_st[lane] = st[lane];
}

// Challenger::get_n_challenges; launch <<<1, 32>>>
__global__ void k_challenger_get(uint64_t* __restrict__ g_st, size_t n, uint64_t* __restrict__ out) {
  __shared__ uint64_t st[32];
  const uint32_t lane = threadIdx.x;
  if (lane < CH_WORDS) st[lane] = g_st[lane];
  __syncwarp();
  for (size_t i = 0; i < n; i++) {
    if (st[CH_NIN] != 0 || st[CH_NOUT] == 0) duplexing_coop(st, lane);
    if (lane == 0) {
      const uint32_t k = (uint32_t)st[CH_NOUT] - 1;
      out[i] = st[CH_OUT + k];
      st[CH_NOUT] = k;
    }
    __syncwarp();
  }
  if (lane < CH_WORDS) g_st[lane] = st[lane];
}

__global__ void k_deinterleave(const uint64_t* __restrict__ in, size_t len, uint64_t* __restrict__ c0,
                               uint64_t* __restrict__ c1) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  ulonglong2 v = reinterpret_cast<const ulonglong2*>(in)[i];
  c0[i] = gl::canon(v.x);
  c1[i] = gl::canon(v.y);
}
__global__ void k_interleave(const uint64_t* __restrict__ c0, const uint64_t* __restrict__ c1, size_t len,
                             uint64_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  reinterpret_cast<ulonglong2*>(out)[i] = make_ulonglong2(c0[i], c1[i]);
}
// out[j] = in[bitrev(j)] for interleaved extension elements (reverse_index_bits_in_place)
__global__ void k_bitrev_ext(const uint64_t* __restrict__ in, uint32_t log_len, uint64_t* __restrict__ out) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >> log_len) return;
  size_t r = log_len ? (size_t)(__brevll((unsigned long long)j) >> (64 - log_len)) : 0;
  ulonglong2 v = reinterpret_cast<const ulonglong2*>(in)[r];
  reinterpret_cast<ulonglong2*>(out)[j] = make_ulonglong2(gl::canon(v.x), gl::canon(v.y));
}

// coeffs'[i] = sum_{j < arity} coeffs[i*arity + j] * beta^j   (Horner from the top coefficient)
__global__ void __launch_bounds__(256)
k_fold_coeffs(const uint64_t* __restrict__ c0, const uint64_t* __restrict__ c1, size_t len, uint32_t arity_bits,
              const uint64_t* __restrict__ beta, uint64_t* __restrict__ o0, uint64_t* __restrict__ o1) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (len >> arity_bits)) return;
  const uint32_t arity = 1u << arity_bits;
  gl::ext2 b{beta[0], beta[1]};
  gl::ext2 acc{0, 0};
  for (uint32_t j = arity; j-- > 0;) {
    acc = gl::ext_mul(acc, b);
    size_t k = (i << arity_bits) + j;
    acc = gl::ext_add(acc, gl::ext2{gl::canon(c0[k]), gl::canon(c1[k])});  // layer 0 reads the caller's raw words
  }
  o0[i] = acc.c0;
  o1[i] = acc.c1;
}

// fri_proof_of_work: the duplex state is the sponge state with the pending inputs written over its head and the
// candidate w at position n_in; response = state[7] after one permutation.  The MINIMAL witness is wanted (the
// reference's rayon find_any is schedule dependent), and it is geometric with mean 2^pow_bits, so a
// fixed chunk wastes most of its permutations.  Every thread walks the candidates base + tid, base + tid + T, ...
// in increasing order and stops as soon as its candidate exceeds the best witness found so far: every candidate
// below the final minimum is still evaluated by its owner (a stale read of `best` only delays a thread's exit),
// and the launch ends one grid stride after the first hit.
__global__ void __launch_bounds__(256)
k_pow_search(const uint64_t* __restrict__ st, uint64_t base, uint64_t count, uint32_t pow_bits,
             unsigned long long* __restrict__ best) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t s0[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s0[i] = st[i];
  const uint32_t n_in = (uint32_t)st[CH_NIN];
#pragma unroll
  for (int i = 0; i < 8; i++)
    if ((uint32_t)i < n_in) s0[i] = st[CH_IN + i];
  for (uint64_t off = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; off < count; off += stride) {
    const uint64_t w = base + off;
    // relaxed GPU-scope load: served by L2, where the atomicMin lands.  (Measured on B200: a plain `volatile` read
    // here kept returning the initial value for the whole launch — every thread ran all its strides.)
    unsigned long long cur;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(best) : "memory");
    if (w >= GL_P || w > cur) return;
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = ((uint32_t)i == n_in) ? w : s0[i];
    // a hit of another thread usually lands while this candidate is in flight (all threads finish a permutation at
    // about the same time): look again a quarter of the way in
    if (!poseidon::permute_nc_abortable(s, [&] {
          unsigned long long now;
          asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(now) : "l"(best) : "memory");
          return w > now;
        }))
      return;
    const uint64_t resp = gl::canon(s[7]);
    if (pow_bits == 0 || (resp >> (64 - pow_bits)) == 0) {
      atomicMin(best, (unsigned long long)w);
      return;
    }
  }
}

}  // namespace frik
