// hash_kernels.cuh — Poseidon leaf hashing and Merkle tree levels (one permutation per thread).
//
// Replaces plonky2 0.2.2 hash/merkle_tree.rs MerkleTree::new (fill_digests_buf / fill_subtree) for
// H = PoseidonHash.  Device layout of a tree: `levels` holds the digests level by level, leaf level
// first: level l (0 = leaf digests) has n_leaves >> l digests of 4 u64, down to the cap level
// (2^cap_height digests).  plonky2's interleaved `digests` vector is produced on demand by
// k_export_plonky2_digests (bit-exact with fill_subtree's layout).
#pragma once
#include "poseidon.cuh"

#ifndef P2B_LEAF_MINB
#define P2B_LEAF_MINB 2  // resident CTAs of 256 threads per SM for the leaf sponge kernels (tuning: profiles/r01_summary.md)
#endif

namespace hashk {

__device__ __forceinline__ void store_digest(uint64_t* __restrict__ dst, const uint64_t d[4]) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(dst);
  p[0] = make_ulonglong2(d[0], d[1]);
  p[1] = make_ulonglong2(d[2], d[3]);
}
__device__ __forceinline__ void load_digest(const uint64_t* __restrict__ src, uint64_t d[4]) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(src);
  ulonglong2 a = p[0], b = p[1];
  d[0] = a.x;
  d[1] = a.y;
  d[2] = b.x;
  d[3] = b.y;
}

// Leaf digests of a column-major matrix: leaf j = (data[c * col_stride + j])_{c < n_cols}.
// hash_or_noop: n_cols <= 4 -> zero-padded copy, else overwrite-mode sponge, 8 columns per permutation.
// Consecutive threads read consecutive j of the same column: every load is a full 256 B per warp.
__global__ void __launch_bounds__(256, P2B_LEAF_MINB)
k_leaf_hash_colmajor(const uint64_t* __restrict__ data, size_t col_stride, uint32_t n_cols, size_t n_leaves,
                     uint64_t* __restrict__ digests) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_leaves) return;
  uint64_t out[4];
  if (n_cols <= 4) {
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = (uint32_t)i < n_cols ? gl::canon(data[(size_t)i * col_stride + j]) : 0;
  } else {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    uint64_t nxt[8];
    const uint64_t* p = data + j;
#pragma unroll
    for (int i = 0; i < 8; i++) nxt[i] = (uint32_t)i < n_cols ? p[(size_t)i * col_stride] : 0;
    for (uint32_t c0 = 0; c0 < n_cols; c0 += 8) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (c0 + i < n_cols) s[i] = nxt[i];
      // prefetch the next 8 columns before the ~20k-instruction permutation
      if (c0 + 8 < n_cols) {
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (c0 + 8 + i < n_cols) nxt[i] = p[(size_t)(c0 + 8 + i) * col_stride];
      }
      poseidon::permute_nc(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
  }
  store_digest(digests + 4 * j, out);
}

// The same sponge in column ranges: absorbs columns [c_begin, c_end) (c_begin a multiple of 8) of every leaf into a
// sponge state kept in HBM between launches (state[i * n_leaves + j], i < 12: coalesced), starting from zero when
// c_begin == 0 and writing the digest instead of the state when c_end == n_cols.  Used by the pipelined host upload
// (p2b.cu batch_from_host_pipelined): leaf hashing of the columns that have arrived overlaps the upload of the rest.
// Only for n_cols > 4 (hash_or_noop's copy case never gets here).
__global__ void __launch_bounds__(256, P2B_LEAF_MINB)
k_leaf_absorb_colmajor(const uint64_t* __restrict__ data, size_t col_stride, uint32_t c_begin, uint32_t c_end,
                       uint32_t n_cols, size_t n_leaves, uint64_t* __restrict__ state, uint64_t* __restrict__ digests) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_leaves) return;
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = c_begin == 0 ? 0 : state[(size_t)i * n_leaves + j];
  uint64_t nxt[8];
  const uint64_t* p = data + j;
#pragma unroll
  for (int i = 0; i < 8; i++) nxt[i] = c_begin + i < c_end ? p[(size_t)(c_begin + i) * col_stride] : 0;
  for (uint32_t c0 = c_begin; c0 < c_end; c0 += 8) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (c0 + i < c_end) s[i] = nxt[i];
    if (c0 + 8 < c_end) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (c0 + 8 + i < c_end) nxt[i] = p[(size_t)(c0 + 8 + i) * col_stride];
    }
    poseidon::permute_nc(s);
  }
  if (c_end == n_cols) {
    uint64_t out[4];
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
    store_digest(digests + 4 * j, out);
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) state[(size_t)i * n_leaves + j] = s[i];
  }
}

// Leaf digests of row-major leaves (MerkleTree::new's own input layout; FRI layer leaves).
__global__ void __launch_bounds__(256)
k_leaf_hash_rowmajor(const uint64_t* __restrict__ leaves, size_t leaf_len, size_t n_leaves,
                     uint64_t* __restrict__ digests) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_leaves) return;
  const uint64_t* p = leaves + j * leaf_len;
  uint64_t out[4];
  if (leaf_len <= 4) {
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = (size_t)i < leaf_len ? gl::canon(p[i]) : 0;
  } else {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (size_t c0 = 0; c0 < leaf_len; c0 += 8) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (c0 + i < leaf_len) s[i] = p[c0 + i];
      poseidon::permute_nc(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
  }
  store_digest(digests + 4 * j, out);
}

// One tree level: parent[i] = two_to_one(child[2i], child[2i+1]).
__global__ void __launch_bounds__(256)
k_tree_level(const uint64_t* __restrict__ child, uint64_t* __restrict__ parent, size_t n_parents) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_parents) return;
  uint64_t l[4], r[4], o[4];
  load_digest(child + 8 * i, l);
  load_digest(child + 8 * i + 4, r);
  poseidon::two_to_one(l, r, o);
  store_digest(parent + 4 * i, o);
}

// The same level for the narrow top of a tree: one warp per node (poseidon::coop_permute_nc), ~4x lower latency
// per level.  blockDim a multiple of 32; grid covers 32 * n_parents threads.
__global__ void __launch_bounds__(256)
k_tree_level_coop(const uint64_t* __restrict__ child, uint64_t* __restrict__ parent, size_t n_parents) {
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t l = threadIdx.x & 31;
  const bool active = g < n_parents;  // idle warps still run (full-mask shuffles inside)
  const size_t gg = active ? g : 0;
  uint64_t s = l < 8 ? child[8 * gg + l] : 0;
  s = poseidon::coop_permute_nc(s, l);
  if (active && l < 4) parent[4 * gg + l] = gl::canon(s);
}

// Row-major leaves, one warp per leaf (FRI layer leaves: few leaves, 4 permutations each)
__global__ void __launch_bounds__(256)
k_leaf_hash_rowmajor_coop(const uint64_t* __restrict__ leaves, size_t leaf_len, size_t n_leaves,
                          uint64_t* __restrict__ digests) {
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t l = threadIdx.x & 31;
  const bool active = g < n_leaves;
  const uint64_t* p = leaves + (active ? g : 0) * leaf_len;
  uint64_t s = 0;
  if (leaf_len <= 4) {
    s = l < leaf_len ? gl::canon(p[l]) : 0;
  } else {
    for (size_t c0 = 0; c0 < leaf_len; c0 += 8) {
      if (l < 8 && c0 + l < leaf_len) s = p[c0 + l];
      if (l >= 12) s = 0;
      s = poseidon::coop_permute_nc(s, l);
    }
    s = gl::canon(s);
  }
  if (active && l < 4) digests[4 * g + l] = s;
}

// n independent pairs (p2b_two_to_one)
__global__ void __launch_bounds__(256)
k_two_to_one_pairs(const uint64_t* __restrict__ left, const uint64_t* __restrict__ right, size_t n,
                   uint64_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t l[4], r[4], o[4];
  load_digest(left + 4 * i, l);
  load_digest(right + 4 * i, r);
  poseidon::two_to_one(l, r, o);
  store_digest(out + 4 * i, o);
}

// n independent permutations, states row-major n x 12
__global__ void __launch_bounds__(256) k_permute_states(uint64_t* __restrict__ states, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = states[12 * i + k];
  poseidon::permute(s);
#pragma unroll
  for (int k = 0; k < 12; k++) states[12 * i + k] = s[k];
}

// levels (leaf level first) -> plonky2's `digests` layout.  L = log2(n_leaves) - cap_height layers are
// stored (the roots live in the cap only).  For the node q of layer i inside cap-subtree t:
//   dst = t * (2^(L+1) - 2) + 2 * (((q >> 1) << (i + 1)) + 2^i - 1) + (q & 1)
__global__ void __launch_bounds__(256)
k_export_plonky2_digests(const uint64_t* __restrict__ levels, size_t n_leaves, uint32_t L,
                         uint64_t* __restrict__ out) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // index over all stored nodes
  size_t total = 2 * n_leaves - (2 * (n_leaves >> L));         // sum_{i<L} n_leaves >> i
  if (g >= total) return;
  // find layer i: nodes of layer i start at offset 2*n_leaves - 2*(n_leaves >> i)
  uint32_t i = 0;
  size_t off = 0;
  while (g >= off + (n_leaves >> i)) {
    off += n_leaves >> i;
    i++;
  }
  size_t node = g - off;              // index inside layer i (whole tree)
  size_t t = node >> (L - i);         // cap subtree
  size_t q = node & ((((size_t)1) << (L - i)) - 1);
  size_t dst = t * ((((size_t)1) << (L + 1)) - 2) + 2 * (((q >> 1) << (i + 1)) + (((size_t)1) << i) - 1) + (q & 1);
  uint64_t d[4];
  load_digest(levels + 4 * g, d);
  store_digest(out + 4 * dst, d);
}

// siblings of `leaf_index` for layers 0..L-1 -> out[L][4]
__global__ void k_gather_proof(const uint64_t* __restrict__ levels, size_t n_leaves, uint32_t L, size_t leaf_index,
                               uint64_t* __restrict__ out) {
  uint32_t i = threadIdx.x >> 2, w = threadIdx.x & 3;
  if (i >= L) return;
  size_t off = 2 * n_leaves - 2 * (n_leaves >> i);
  size_t sib = (leaf_index >> i) ^ 1;
  out[4 * i + w] = levels[4 * (off + sib) + w];
}

// row `j` of a column-major matrix -> out[n_cols]
__global__ void k_gather_row_colmajor(const uint64_t* __restrict__ data, size_t col_stride, uint32_t n_cols, size_t j,
                                      uint64_t* __restrict__ out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cols) out[c] = data[(size_t)c * col_stride + j];
}

// column-major (n_cols x n_rows) -> row-major (n_rows x n_cols) through a 32x32 shared tile
__global__ void __launch_bounds__(256)
k_transpose_to_rowmajor(const uint64_t* __restrict__ in, size_t col_stride, uint32_t n_cols, size_t n_rows,
                        uint64_t* __restrict__ out) {
  __shared__ uint64_t tile[32][33];
  size_t r0 = (size_t)blockIdx.x * 32;
  uint32_t c0 = blockIdx.y * 32;
  uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (uint32_t k = ty; k < 32; k += 8) {
    uint32_t c = c0 + k;
    size_t r = r0 + tx;
    if (c < n_cols && r < n_rows) tile[k][tx] = in[(size_t)c * col_stride + r];
  }
  __syncthreads();
  for (uint32_t k = ty; k < 32; k += 8) {
    size_t r = r0 + k;
    uint32_t c = c0 + tx;
    if (c < n_cols && r < n_rows) out[r * n_cols + c] = tile[tx][k];
  }
}

// hash_no_pad of one vector (public inputs hash, known-answer tests): one warp-cooperative sponge; <<<1, 32>>>
__global__ void k_hash_no_pad_single(const uint64_t* __restrict__ in, size_t len, uint64_t* __restrict__ out) {
  const uint32_t l = threadIdx.x & 31;
  uint64_t s = 0;
  for (size_t c0 = 0; c0 < len; c0 += 8) {
    if (l < 8 && c0 + l < len) s = in[c0 + l];
    if (l >= 12) s = 0;
    s = poseidon::coop_permute_nc(s, l);
  }
  if (threadIdx.x < 4) out[threadIdx.x] = gl::canon(s);
}

}  // namespace hashk
