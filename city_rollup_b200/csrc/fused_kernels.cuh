// fused_kernels.cuh — the small-proof path of p2b_prove: fewer, fused launches and no host round trip.
//
// A City Rollup worker proof is tiny for a B200 (2^12 rows: SURVEY.md §0.6), so its cost is the LENGTH of the
// dependent launch chain, not arithmetic.  These kernels replace chains of one-purpose launches of the stage-by-stage
// path (the stage entry points of include/p2b.h keep theirs):
//   k_tree_subtree      all levels of a Merkle tree in ONE launch: every CTA climbs its 512-digest subtree in shared
//                       memory, the last CTA to finish climbs the top (plonky2 hash/merkle_tree.rs fill_digests_buf;
//                       was one launch per level)
//   k_transcript        Challenger: observe a list of device segments, squeeze, optional derived values
//                       (iop/challenger.rs observe_elements / get_n_challenges; was one launch per call)
//   k_pow_finish        fri_proof_of_work's tail and the query indices, fed by the device-side witness search
//   k_eval_polys_multi  OpeningSet::new: every opened polynomial in one launch (was one per field of the set)
//   k_query_all         fri_prover_query_rounds: every leaf row and Merkle path of every query in one launch
//   k_reduce_polys2 / k_divide_by_linear2 / k_final_poly_combine   PolynomialBatch::prove_openings' final polynomial
//   k_leaf_hash_planes[_coop] FRI layer leaves hashed straight from the two value planes (one thread / one warp per leaf)
// Every result is bit-identical to the unfused kernels (same field operations on the same operands).
#pragma once
#include "fri_kernels.cuh"
#include "hash_kernels.cuh"
#include "prover_kernels.cuh"

namespace fusedk {

using gl::ext2;

// ------------------------------------------------------------------------------------------------ Merkle tree
// levels: the tree's digest array, leaf level first (hash_kernels.cuh); level `first` (n_first = 2^log_first digests)
// is already there.  Grid = max(1, n_first >> 9) CTAs of 256 threads; L_rem = number of levels still to build
// above level `first` (down to the cap).  CTA b climbs the subtree over digests [b * 512, (b + 1) * 512) of level
// `first` in shared memory: one thread per node while a level has more than COOP_NODES nodes in the CTA (a
// full-throughput permutation per thread), one warp per node for the last few (the warp-cooperative permutation: ~1/3
// of the latency).  The CTA that finishes last (device-scope counter, reset before it leaves so that the launch can be
// replayed from a CUDA graph) climbs what is left above the subtree roots the same way.
constexpr int SUB_LOG = 9;
// Levels with at most this many nodes in a CTA go to one warp per node.  A warp-cooperative permutation finishes in a
// third of the time of a one-thread permutation but executes ~10x the instructions (5.5k warp instructions against
// 16.5k thread instructions = 516 warp-instruction equivalents).  The fused kernel only sees the last 2048 digests of
// a tree (p2b.cu build_levels), so its cooperative nodes are a negligible share of a proof's instructions and the
// threshold is set for latency: 16 (two permutations per warp: 15 us against 23).
constexpr uint32_t COOP_NODES = 16;

__device__ __forceinline__ void climb_in_smem(uint64_t (*sm)[4], uint32_t n_in, uint32_t n_levels, uint64_t* __restrict__ levels,
                                              size_t n_leaves, uint32_t first_level, size_t node0) {
  // sm[0 .. n_in) holds the digests of the current level; node0 = index of sm[0]'s PARENT-level sibling group base
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  uint32_t cur = n_in;
  size_t base = node0;  // global index (within its level) of sm[0]
  for (uint32_t l = 0; l < n_levels; l++) {
    const uint32_t n_par = cur >> 1;
    base >>= 1;
    uint64_t* out = levels + 4 * (2 * n_leaves - 2 * (n_leaves >> (first_level + l + 1)));
    if (n_par > COOP_NODES) {
      uint64_t o[4];
      const bool act = tid < n_par;
      if (act) poseidon::two_to_one(sm[2 * tid], sm[2 * tid + 1], o);
      __syncthreads();
      if (act) {
#pragma unroll
        for (int k = 0; k < 4; k++) sm[tid][k] = o[k];
        hashk::store_digest(out + 4 * (base + tid), o);
      }
      __syncthreads();
    } else {
      // warp-cooperative: node w, w + n_warps, ...  (n_par <= COOP_NODES <= 2 n_warps: at most two nodes per warp)
      uint64_t res[2];
      uint32_t cnt = 0;
      for (uint32_t node = warp; node < n_par; node += n_warps) {
        uint64_t s = lane < 8 ? sm[2 * node + (lane >> 2)][lane & 3] : 0;
        s = gl::canon(poseidon::coop_permute_nc(s, lane));
        res[cnt++] = s;
      }
      __syncthreads();
      cnt = 0;
      for (uint32_t node = warp; node < n_par; node += n_warps) {
        if (lane < 4) {
          sm[node][lane] = res[cnt];
          out[4 * (base + node) + lane] = res[cnt];
        }
        cnt++;
      }
      __syncthreads();
    }
    cur = n_par;
  }
}

#ifndef P2B_TREE_MINB
#define P2B_TREE_MINB 1
#endif
__global__ void __launch_bounds__(256, P2B_TREE_MINB)
k_tree_subtree(uint64_t* __restrict__ levels, size_t n_leaves, uint32_t first_level, uint32_t log_first, uint32_t L_rem,
               unsigned int* __restrict__ counter) {
  __shared__ uint64_t sm[1 << SUB_LOG][4];
  __shared__ bool is_last;
  const uint32_t tid = threadIdx.x;
  const uint32_t sub_log = log_first < (uint32_t)SUB_LOG ? log_first : (uint32_t)SUB_LOG;
  const uint32_t n_in = 1u << sub_log;
  const uint32_t lv_here = L_rem < sub_log ? L_rem : sub_log;
  const uint64_t* src = levels + 4 * (2 * n_leaves - 2 * (n_leaves >> first_level)) + 4 * ((size_t)blockIdx.x << sub_log);
  for (uint32_t i = tid; i < n_in; i += blockDim.x) hashk::load_digest(src + 4 * i, sm[i]);
  __syncthreads();
  climb_in_smem(sm, n_in, lv_here, levels, n_leaves, first_level, (size_t)blockIdx.x << sub_log);
  if (L_rem <= sub_log) return;
  // hand the subtree roots to the last CTA
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) *counter = 0;  // replayable: nobody else touches it any more
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the level of subtree roots: gridDim.x = 2^(log_first - sub_log) digests, at level first + sub_log
  uint32_t lvl = first_level + sub_log, log_cur = log_first - sub_log, rem = L_rem - sub_log;
  while (rem > 0) {
    const uint32_t n_cur = 1u << log_cur;
    // at most 512 digests at a time through shared memory; a wider top is walked in 512-digest pieces level by level
    const uint32_t piece_log = log_cur < (uint32_t)SUB_LOG ? log_cur : (uint32_t)SUB_LOG;
    const uint32_t pieces = n_cur >> piece_log;
    const uint32_t lv = rem < piece_log ? rem : piece_log;
    for (uint32_t pc = 0; pc < pieces; pc++) {
      const uint64_t* s2 = levels + 4 * (2 * n_leaves - 2 * (n_leaves >> lvl)) + 4 * ((size_t)pc << piece_log);
      for (uint32_t i = tid; i < (1u << piece_log); i += blockDim.x) {
        // written by other CTAs in this launch: read through L2
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(s2 + 4 * i);
        const ulonglong2 a = __ldcg(p), b = __ldcg(p + 1);
        sm[i][0] = a.x, sm[i][1] = a.y, sm[i][2] = b.x, sm[i][3] = b.y;
      }
      __syncthreads();
      climb_in_smem(sm, 1u << piece_log, lv, levels, n_leaves, lvl, (size_t)pc << piece_log);
    }
    lvl += lv;
    log_cur -= lv;
    rem -= lv;
  }
}

// FRI layer leaves: leaf j = (c0[j * arity + k], c1[j * arity + k])_k, read straight from the two planes (leaf order),
// hashed by one warp per leaf, and written row-major for the query rounds (MerkleTree::get).
__global__ void __launch_bounds__(256)
k_leaf_hash_planes_coop(const uint64_t* __restrict__ c0, const uint64_t* __restrict__ c1, uint32_t arity_bits, size_t n_leaves,
                        uint64_t* __restrict__ leaves_rm, uint64_t* __restrict__ digests) {
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t l = threadIdx.x & 31;
  const bool active = g < n_leaves;
  const size_t gg = active ? g : 0;
  const uint32_t leaf_len = 2u << arity_bits;
  const uint64_t* p0 = c0 + (gg << arity_bits);
  const uint64_t* p1 = c1 + (gg << arity_bits);
  uint64_t* row = leaves_rm + gg * leaf_len;
  uint64_t s = 0;
  if (leaf_len <= 4) {  // arity 2: hash_or_noop copies
    uint64_t v = 0;
    if (l < leaf_len) v = gl::canon((l & 1) ? p1[l >> 1] : p0[l >> 1]);
    if (active && l < leaf_len) row[l] = v;
    s = v;
  } else {
    for (uint32_t q0 = 0; q0 < leaf_len; q0 += 8) {
      if (l < 8 && q0 + l < leaf_len) {
        const uint32_t q = q0 + l;
        s = gl::canon((q & 1) ? p1[q >> 1] : p0[q >> 1]);
        if (active) row[q] = s;
      }
      if (l >= 12) s = 0;
      s = poseidon::coop_permute_nc(s, l);
    }
    s = gl::canon(s);
  }
  if (active && l < 4) digests[4 * g + l] = s;
}

// The same with one thread per leaf (the full-throughput permutation): layers with thousands of leaves.
__global__ void __launch_bounds__(256)
k_leaf_hash_planes(const uint64_t* __restrict__ c0, const uint64_t* __restrict__ c1, uint32_t arity_bits, size_t n_leaves,
                   uint64_t* __restrict__ leaves_rm, uint64_t* __restrict__ digests) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_leaves) return;
  const uint32_t arity = 1u << arity_bits, leaf_len = 2 * arity;
  const uint64_t* p0 = c0 + (g << arity_bits);
  const uint64_t* p1 = c1 + (g << arity_bits);
  ulonglong2* row = reinterpret_cast<ulonglong2*>(leaves_rm + g * leaf_len);
  uint64_t out[4];
  if (leaf_len <= 4) {  // arity 2: hash_or_noop copies
    const uint64_t a = gl::canon(p0[0]), b = gl::canon(p1[0]), c = gl::canon(p0[1]), d = gl::canon(p1[1]);
    row[0] = make_ulonglong2(a, b);
    row[1] = make_ulonglong2(c, d);
    out[0] = a, out[1] = b, out[2] = c, out[3] = d;
  } else {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (uint32_t e0 = 0; e0 < arity; e0 += 4) {  // 4 extension elements = 8 words per permutation
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint64_t a = gl::canon(p0[e0 + k]), b = gl::canon(p1[e0 + k]);
        row[e0 + k] = make_ulonglong2(a, b);
        s[2 * k] = a;
        s[2 * k + 1] = b;
      }
      poseidon::permute_nc(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = gl::canon(s[i]);
  }
  hashk::store_digest(digests + 4 * g, out);
}

// ------------------------------------------------------------------------------------------------ transcript
constexpr int TR_MAX_SEGS = 6;
struct TranscriptParams {
  uint64_t* state;  // frik::ChallengerState
  const uint64_t* seg[TR_MAX_SEGS];
  uint32_t seg_len[TR_MAX_SEGS];
  uint32_t n_seg;
  uint32_t reset;    // start from Challenger::new()
  uint64_t* out;     // n_out squeezed challenges
  uint32_t n_out;
  // derived values, computed by the same warp once the challenges exist:
  //  scale_g != 0: out[2..4) = g * (out[0], out[1])                 (zeta_next = g * zeta)
  //  pow_tab != nullptr: base-field power tables pow_tab[c * pow_n + k] = out[c]^k, c < n_out   (alpha powers of the
  //                      quotient's reduce_with_powers)
  //  ext_tab != nullptr: ext_tab[2 k .. 2 k + 2) = (out[0], out[1])^k as an extension element, k <= ext_n
  uint64_t scale_g;
  // subgroup_check_bits = degree_bits + 1 (0 = no check): *subgroup_flag = 1 when (out[0], out[1])^(2^degree_bits) == 1,
  // i.e. the opening point zeta lies in the subgroup H (plonky2's prover refuses: "Opening point is in the subgroup"), else 0
  uint32_t subgroup_check_bits;
  uint64_t* subgroup_flag;
  uint64_t* pow_tab;
  uint32_t pow_n;
  uint64_t* ext_tab;
  uint32_t ext_n;
};

__device__ __forceinline__ void ch_observe(uint64_t* st, uint32_t lane, const uint64_t* __restrict__ elems, size_t n) {
  size_t i = 0;
  while (i < n) {
    const uint32_t k = (uint32_t)st[frik::CH_NIN];
    const uint32_t take = (uint32_t)((n - i) < (size_t)(8 - k) ? (n - i) : (size_t)(8 - k));
    __syncwarp();
    if (lane < take) st[frik::CH_IN + k + lane] = gl::canon(elems[i + lane]);
    if (lane == 0) {
      st[frik::CH_NOUT] = 0;
      st[frik::CH_NIN] = k + take;
    }
    __syncwarp();
    i += take;
    if (k + take == 8) frik::duplexing_coop(st, lane);
  }
  __syncwarp();
}
__device__ __forceinline__ void ch_get(uint64_t* st, uint32_t lane, size_t n, uint64_t* __restrict__ out) {
  for (size_t i = 0; i < n; i++) {
    if (st[frik::CH_NIN] != 0 || st[frik::CH_NOUT] == 0) frik::duplexing_coop(st, lane);
    if (lane == 0) {
      const uint32_t k = (uint32_t)st[frik::CH_NOUT] - 1;
      out[i] = st[frik::CH_OUT + k];
      st[frik::CH_NOUT] = k;
    }
    __syncwarp();
  }
}

// <<<1, 32>>>
__global__ void k_transcript(TranscriptParams P) {
  __shared__ uint64_t st[32];
  __shared__ uint64_t chal[8];
  const uint32_t lane = threadIdx.x;
  if (lane < frik::CH_WORDS) st[lane] = P.reset ? 0 : P.state[lane];
  __syncwarp();
  for (uint32_t s = 0; s < P.n_seg; s++) ch_observe(st, lane, P.seg[s], P.seg_len[s]);
  if (P.n_out) {
    // squeeze into shared memory first (<= 8 at a time is all the derived values need), then to `out`
    for (uint32_t i0 = 0; i0 < P.n_out; i0 += 8) {
      const uint32_t k = P.n_out - i0 < 8 ? P.n_out - i0 : 8;
      ch_get(st, lane, k, chal);
      __syncwarp();
      if (lane < k) P.out[i0 + lane] = chal[lane];
      __syncwarp();
    }
  }
  if (lane < frik::CH_WORDS) P.state[lane] = st[lane];
  if (P.n_out == 0 || P.n_out > 8) return;
  if (P.scale_g && lane < 2) P.out[2 + lane] = gl::mul(chal[lane], P.scale_g);
  if (P.subgroup_check_bits && lane == 0) {
    ext2 z{chal[0], chal[1]};
    for (uint32_t k = 1; k < P.subgroup_check_bits; k++) z = gl::ext_mul(z, z);
    *P.subgroup_flag = (z.c0 == 1 && z.c1 == 0) ? 1 : 0;
  }
  if (P.pow_tab) {
    // lane l fills entries [l * per, (l + 1) * per) of every table
    const uint32_t per = (P.pow_n + 31) / 32;
    for (uint32_t c = 0; c < P.n_out; c++) {
      const uint64_t a = chal[c];
      uint64_t p = gl::pow(a, (uint64_t)lane * per);
      for (uint32_t k = lane * per; k < (lane + 1) * per && k < P.pow_n; k++) {
        P.pow_tab[(size_t)c * P.pow_n + k] = p;
        p = gl::mul(p, a);
      }
    }
  }
  if (P.ext_tab) {
    const ext2 a{chal[0], chal[1]};
    const uint32_t cnt = P.ext_n + 1, per = (cnt + 31) / 32;
    ext2 p = provk::ext_pow(a, (size_t)lane * per);
    for (uint32_t k = lane * per; k < (lane + 1) * per && k < cnt; k++) {
      P.ext_tab[2 * k] = p.c0;
      P.ext_tab[2 * k + 1] = p.c1;
      p = gl::ext_mul(p, a);
    }
  }
}

// fri_proof_of_work's tail: challenger.observe_element(w); pow_response = challenger.get_challenge(); then the
// query indices' challenges (fri_prover_query_rounds squeezes one per round).  `best` is what k_pow_search left
// (the minimal witness); it is copied to the proof and reset for the next search.  <<<1, 32>>>
__global__ void k_pow_finish(uint64_t* __restrict__ g_st, unsigned long long* __restrict__ best, uint64_t* __restrict__ witness_out,
                             uint32_t n_queries, uint64_t* __restrict__ chal_out) {
  __shared__ uint64_t st[32];
  __shared__ uint64_t tmp[2];
  const uint32_t lane = threadIdx.x;
  if (lane < frik::CH_WORDS) st[lane] = g_st[lane];
  if (lane == 0) tmp[0] = *best;
  __syncwarp();
  ch_observe(st, lane, tmp, 1);
  ch_get(st, lane, 1, tmp + 1);
  ch_get(st, lane, n_queries, chal_out);
  if (lane < frik::CH_WORDS) g_st[lane] = st[lane];
  if (lane == 0) {
    *witness_out = tmp[0];
    *best = ~0ull;
  }
}

// ------------------------------------------------------------------------------------------------ openings
constexpr int EV_MAX_RANGES = 8;
struct EvalParams {
  const uint64_t* coeffs[EV_MAX_RANGES];  // first polynomial of the range (column-major, n coefficients each)
  uint64_t* out[EV_MAX_RANGES];           // 2 words per polynomial
  uint32_t first_cta[EV_MAX_RANGES + 1];  // prefix sums of the polynomial counts
  uint32_t point[EV_MAX_RANGES];          // which of the points (2 words each) the range is opened at
  uint32_t n_ranges;
  size_t n;
  const uint64_t* points;
};
// one CTA per polynomial, all ranges in one launch (the body of provk::k_eval_polys_ext)
__global__ void __launch_bounds__(256) k_eval_polys_multi(EvalParams P) {
  __shared__ uint64_t s0[256], s1[256];
  uint32_t r = 0;
  while (r + 1 < P.n_ranges && blockIdx.x >= P.first_cta[r + 1]) r++;
  const uint32_t idx = blockIdx.x - P.first_cta[r];
  const size_t n = P.n;
  const uint64_t* c = P.coeffs[r] + (size_t)idx * n;
  const uint64_t* zp = P.points + 2 * P.point[r];
  __shared__ ext2 tab[48];
  const ext2 z{gl::canon(zp[0]), gl::canon(zp[1])};
  const ext2 val = provk::eval_poly_cta(c, n, z, s0, s1, tab);
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    P.out[r][2 * idx] = val.c0;
    P.out[r][2 * idx + 1] = val.c1;
  }
}

// ------------------------------------------------------------------------------------------------ prove_openings
// alpha.reduce_polys_base for both opening batches in one launch.  Batch 0 (everything at zeta) is split into
// `slices` slices of consecutive polynomials, slice y < slices writing its partial sum to part[y] (two planes of n:
// field addition is exact, so summing the partial sums later gives the same element); row y == slices computes batch 1
// (m1 polynomials: the Zs again at g * zeta) completely.  pw[2 i ..] = alpha^i.
__global__ void __launch_bounds__(256)
k_reduce_polys2(const uint64_t* const* __restrict__ polys0, uint32_t m0, const uint64_t* const* __restrict__ polys1, uint32_t m1,
                size_t n, const uint64_t* __restrict__ pw, uint32_t slices, uint64_t* __restrict__ part, uint64_t* __restrict__ comp1) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t y = blockIdx.y;
  const uint64_t* const* polys = y < slices ? polys0 : polys1;
  const uint32_t per = (m0 + slices - 1) / slices;
  const uint32_t lo = y < slices ? y * per : 0, hi = y < slices ? (lo + per < m0 ? lo + per : m0) : m1;
  uint64_t a0 = 0, a1 = 0;
  for (uint32_t i = lo; i < hi; i++) {
    const uint64_t c = gl::canon(polys[i][k]);
    a0 = gl::add(a0, gl::mul(c, pw[2 * i]));
    a1 = gl::add(a1, gl::mul(c, pw[2 * i + 1]));
  }
  uint64_t* o = y < slices ? part + (size_t)y * 2 * n : comp1;
  o[k] = a0;
  o[n + k] = a1;
}

// divide_by_linear for both batches: CTA 0 sums the partial sums of batch 0 and divides by (X - zeta), CTA 1 divides
// batch 1 by (X - g zeta); the body is provk::k_divide_by_linear's.  points = zeta | zeta_next.
__global__ void __launch_bounds__(1024)
k_divide_by_linear2(const uint64_t* __restrict__ part, uint32_t slices, const uint64_t* __restrict__ comp1, size_t n,
                    const uint64_t* __restrict__ points, uint64_t* __restrict__ quot /* [2 batches][2 planes][n] */,
                    uint64_t* __restrict__ comp0_scratch /* [2][n] */) {
  __shared__ uint64_t L0[1024], L1[1024], Z0[1024], Z1[1024];
  const uint32_t tid = threadIdx.x, b = blockIdx.x;
  const uint64_t *c0, *c1;
  if (b == 0) {
    for (size_t k = tid; k < n; k += 1024) {
      uint64_t a0 = 0, a1 = 0;
      for (uint32_t y = 0; y < slices; y++) {
        a0 = gl::add(a0, part[(size_t)y * 2 * n + k]);
        a1 = gl::add(a1, part[(size_t)y * 2 * n + n + k]);
      }
      comp0_scratch[k] = a0;
      comp0_scratch[n + k] = a1;
    }
    __syncthreads();
    c0 = comp0_scratch, c1 = comp0_scratch + n;
  } else {
    c0 = comp1, c1 = comp1 + n;
  }
  uint64_t* q0 = quot + (size_t)b * 2 * n;
  uint64_t* q1 = q0 + n;
  const uint64_t* zptr = points + 2 * b;
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
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    ext2 l{L0[tid], L1[tid]}, p{Z0[tid], Z1[tid]};
    const bool has = tid + off < 1024;
    ext2 l2{0, 0}, p2{1, 0};
    if (has) l2 = ext2{L0[tid + off], L1[tid + off]}, p2 = ext2{Z0[tid + off], Z1[tid + off]};
    __syncthreads();
    if (has) {
      l = gl::ext_add(l, gl::ext_mul(p, l2));
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

// final_poly = quot0 * alpha^m1 + quot1 (alpha.shift_poly between the batches), written twice: as the planes the LDE
// reads (fin, n each) and zero-padded to N as the FRI coefficient planes (coef, N each).  s = alpha^m1.
__global__ void __launch_bounds__(256)
k_final_poly_combine(const uint64_t* __restrict__ quot, size_t n, size_t N, const uint64_t* __restrict__ s, uint64_t* __restrict__ fin,
                     uint64_t* __restrict__ coef) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  uint64_t r0 = 0, r1 = 0;
  if (k < n) {
    const ext2 r = gl::ext_mul(ext2{quot[k], quot[n + k]}, ext2{s[0], s[1]});
    r0 = gl::add(r.c0, quot[2 * n + k]);
    r1 = gl::add(r.c1, quot[3 * n + k]);
    fin[k] = r0;
    fin[n + k] = r1;
  }
  coef[k] = r0;
  coef[N + k] = r1;
}

// ------------------------------------------------------------------------------------------------ query rounds
constexpr int Q_MAX_ORACLES = 4, Q_MAX_LAYERS = 16;
struct QueryParams {
  // initial trees (column-major batches)
  const uint64_t* o_data[Q_MAX_ORACLES];
  const uint64_t* o_levels[Q_MAX_ORACLES];
  uint32_t o_cols[Q_MAX_ORACLES], o_L[Q_MAX_ORACLES], o_off[Q_MAX_ORACLES];  // o_off: word offset inside a query round
  // commit-phase layers (row-major leaves)
  const uint64_t* l_leaves[Q_MAX_LAYERS];
  const uint64_t* l_levels[Q_MAX_LAYERS];
  uint32_t l_leaf_len[Q_MAX_LAYERS], l_L[Q_MAX_LAYERS], l_shift[Q_MAX_LAYERS], l_off[Q_MAX_LAYERS], l_log_leaves[Q_MAX_LAYERS];
  uint32_t n_oracles, n_layers, log_lde;
  size_t N, per_query;
  const uint64_t* chal;
  uint64_t* out;
};
// grid (n_queries, n_oracles + n_layers): MerkleTree::get + MerkleTree::prove of one tree for one query
__global__ void __launch_bounds__(128) k_query_all(QueryParams P) {
  const uint32_t q = blockIdx.x, item = blockIdx.y;
  uint64_t* out = P.out + (size_t)q * P.per_query;
  if (item < P.n_oracles) {
    const size_t x = provk::query_index(P.chal, q, P.log_lde, 0);
    const uint32_t nc = P.o_cols[item], L = P.o_L[item];
    uint64_t* o = out + P.o_off[item];
    for (uint32_t c = threadIdx.x; c < nc; c += blockDim.x) o[c] = P.o_data[item][(size_t)c * P.N + x];
    o += nc;
    for (uint32_t t = threadIdx.x; t < 4 * L; t += blockDim.x) {
      const uint32_t i = t >> 2, w = t & 3;
      const size_t off = 2 * P.N - 2 * (P.N >> i);
      o[t] = P.o_levels[item][4 * (off + ((x >> i) ^ 1)) + w];
    }
  } else {
    const uint32_t l = item - P.n_oracles;
    const size_t x = provk::query_index(P.chal, q, P.log_lde, P.l_shift[l]);
    const uint32_t len = P.l_leaf_len[l], L = P.l_L[l];
    const size_t n_leaves = (size_t)1 << P.l_log_leaves[l];
    uint64_t* o = out + P.l_off[l];
    for (uint32_t c = threadIdx.x; c < len; c += blockDim.x) o[c] = P.l_leaves[l][x * len + c];
    o += len;
    for (uint32_t t = threadIdx.x; t < 4 * L; t += blockDim.x) {
      const uint32_t i = t >> 2, w = t & 3;
      const size_t off = 2 * n_leaves - 2 * (n_leaves >> i);
      o[t] = P.l_levels[l][4 * (off + ((x >> i) ^ 1)) + w];
    }
  }
}

// device-to-device gather of small pieces into the proof buffer (caps, final polynomial, public inputs): one launch
constexpr int CP_MAX = 24;
struct CopyParams {
  const uint64_t* src[CP_MAX];
  uint64_t* dst[CP_MAX];
  uint32_t len[CP_MAX];
  uint32_t n;
};
__global__ void __launch_bounds__(256) k_copy_multi(CopyParams P) {
  const uint32_t b = blockIdx.x;
  if (b >= P.n) return;
  for (uint32_t i = threadIdx.x; i < P.len[b]; i += blockDim.x) P.dst[b][i] = P.src[b][i];
}

}  // namespace fusedk
