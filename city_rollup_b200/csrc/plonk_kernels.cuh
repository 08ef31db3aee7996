// plonk_kernels.cuh — the PLONK-side prover stages between the three commitments, on device-resident batches.
//
// Replaces, for F = Goldilocks / D = 2 / no lookups / no blinding (every City Rollup worker circuit,
// SURVEY.md §8(c)), plonky2 0.2.2:
//   plonk/prover.rs          wires_permutation_partial_products_and_zs  (k_pp_*)
//   plonk/prover.rs          compute_quotient_polys                     (k_quotient + coset iNTT in p2b.cu)
//   plonk/vanishing_poly.rs  eval_vanishing_poly_base_batch, evaluate_gate_constraints_base_batch
//   plonk/plonk_common.rs    ZeroPolyOnCoset::{eval, eval_inverse, eval_l_0}, reduce_with_powers_multi
//   gates/gate.rs            eval_filtered_base_batch / compute_filter
//   gates/{noop,constant,public_input,arithmetic_base,base_sum,poseidon,arithmetic_extension,
//          multiplication_extension,reducing,reducing_extension,random_access,poseidon_mds,
//          coset_interpolation}.rs  eval_unfiltered_base_*
// and the in-tree gates of the reference (scalar eval_unfiltered is the specification):
//   city_common_circuit/src/u32/gates/arithmetic_u32.rs:88-150, add_many_u32.rs:87-135,
//   subtraction_u32.rs:82-125, range_check_u32.rs:51-75, interleave_u32.rs:86-127, uninterleave_to_u32.rs:93-136,
//   uninterleave_to_b32.rs:97-141, comparison.rs:96-170.
//
// Design.  One thread = one point of the quotient LDE coset.  plonky2 materialises every constraint of
// every gate for a batch of 32 points and then reduces them with the powers of alpha; here every constraint
// is multiplied by its power of alpha and accumulated the moment it is produced (exact arithmetic, so the
// result is identical), which needs no per-point constraint storage at all.  Threads walk the points in
// LEAF order, so that all loads from the three column-major leaf-ordered LDE matrices are coalesced;
// only the two "next row" Z values and the quotient store are scattered.
#pragma once
#include "gl64.cuh"
#include "ntt2_kernels.cuh"
#include "poseidon.cuh"

namespace plonk {

#define PFAST_QUAL __device__
#include "poseidon_fast.inc"
#undef PFAST_QUAL

enum GateKind : uint32_t {
  GATE_NOOP = 0,
  GATE_CONSTANT = 1,
  GATE_PUBLIC_INPUT = 2,
  GATE_ARITHMETIC = 3,
  GATE_POSEIDON = 4,
  GATE_BASE_SUM = 5,
  GATE_U32_ARITHMETIC = 6,
  GATE_U32_ADD_MANY = 7,
  GATE_U32_SUBTRACTION = 8,
  GATE_U32_RANGE_CHECK = 9,
  GATE_U32_INTERLEAVE = 10,
  GATE_UNINTERLEAVE_TO_U32 = 11,
  GATE_UNINTERLEAVE_TO_B32 = 12,
  GATE_COMPARISON = 13,
  GATE_ARITHMETIC_EXT = 14,
  GATE_MUL_EXT = 15,
  GATE_REDUCING = 16,
  GATE_REDUCING_EXT = 17,
  GATE_RANDOM_ACCESS = 18,
  GATE_POSEIDON_MDS = 19,
  GATE_COSET_INTERPOLATION = 20,
  GATE_KIND_COUNT = 21
};

struct Gate {  // mirrors p2b_gate (include/p2b.h)
  uint32_t kind, p0, p1, selector_index, group_start, group_end, row;
};

constexpr int MAX_CHALLENGES = 4;
#define P2B_UNUSED_SELECTOR 0xFFFFFFFFull

__device__ __forceinline__ uint64_t fadd(uint64_t a, uint64_t b) { return gl::add(a, b); }
__device__ __forceinline__ uint64_t fsub(uint64_t a, uint64_t b) { return gl::sub(a, b); }
__device__ __forceinline__ uint64_t fmul(uint64_t a, uint64_t b) { return gl::mul(a, b); }

// ------------------------------------------------------------------------------------------ partial products / Z
struct PpParams {
  const uint64_t* wires;   // num_wires x n values on H (column-major)
  const uint64_t* sigmas;  // num_routed x n values on H
  const uint64_t* k_is;
  uint64_t* local;         // [challenge][chunk][n]: prefix products of the row's chunk quotients
  const uint64_t* betas;   // device: the transcript's challenges never have to visit the host
  const uint64_t* gammas;
  uint32_t log_n, num_routed, chunk, n_chunks;
  ntt2::RootTables roots;
};

// grid (n / 256, num_challenges): per row, the quotient of every chunk of `chunk` routed wires
// prod(w + beta k_j x + gamma) / prod(w + beta sigma_j + gamma), as running products over the chunks
__global__ void __launch_bounds__(256) k_pp_rows(PpParams P) {
  const size_t n = (size_t)1 << P.log_n;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ch = blockIdx.y;
  const uint64_t beta = gl::canon(P.betas[ch]), gamma = gl::canon(P.gammas[ch]);
  const uint64_t bx = fmul(beta, ntt2::root_pow(P.roots, P.log_n, i));  // beta * w^i
  uint64_t* out = P.local + (size_t)ch * P.n_chunks * n + i;
  // The chunk quotients np_k / dp_k need an inversion each (~2.8k instructions through Fermat); with up to
  // PP_MAX_CHUNKS chunks their denominators are inverted together (Montgomery's trick: one inversion and three
  // multiplications per chunk) — the inverse of a field element is unique, so the quotients are the same elements.
  constexpr uint32_t PP_MAX_CHUNKS = 16;
  if (P.n_chunks <= PP_MAX_CHUNKS) {
    uint64_t np[PP_MAX_CHUNKS], pre[PP_MAX_CHUNKS];  // pre[k] = dp_0 * ... * dp_k
    uint64_t run = 1;
#pragma unroll 1
    for (uint32_t k = 0; k < P.n_chunks; k++) {
      uint64_t num_p = 1, dp = 1;
      for (uint32_t j = k * P.chunk; j < (k + 1) * P.chunk && j < P.num_routed; j++) {
        const uint64_t wg = fadd(gl::canon(P.wires[(size_t)j * n + i]), gamma);
        num_p = gl::mul_nc(num_p, gl::mad_nc(bx, P.k_is[j], wg));
        dp = gl::mul_nc(dp, gl::mad_nc(beta, P.sigmas[(size_t)j * n + i], wg));
      }
      num_p = gl::canon(num_p);
      dp = gl::canon(dp);
      np[k] = num_p;
      // a zero denominator makes plonky2's batch inversion panic; here it would poison the whole row: keep the chunk's
      // own inverse semantics (0^-1 = 0 through Fermat) by leaving zero factors out of the running product
      pre[k] = dp;
      run = dp ? fmul(run, dp) : run;
    }
    uint64_t inv_run = gl::inv(run);  // 1 / product of the non-zero denominators
    // walk back: inv(dp_k) = inv_run * (product of the non-zero dp_j, j < k), then drop dp_k from inv_run
    uint64_t q[PP_MAX_CHUNKS];
    // prefix products of the non-zero denominators
    uint64_t pref[PP_MAX_CHUNKS];
    uint64_t pp = 1;
#pragma unroll 1
    for (uint32_t k = 0; k < P.n_chunks; k++) {
      pref[k] = pp;
      if (pre[k]) pp = fmul(pp, pre[k]);
    }
#pragma unroll 1
    for (uint32_t k = P.n_chunks; k-- > 0;) {
      if (pre[k]) {
        const uint64_t inv_k = fmul(inv_run, pref[k]);
        inv_run = fmul(inv_run, pre[k]);
        q[k] = fmul(np[k], inv_k);
      } else {
        q[k] = 0;
      }
    }
    uint64_t acc = 1;
#pragma unroll 1
    for (uint32_t k = 0; k < P.n_chunks; k++) {
      acc = fmul(acc, q[k]);
      out[(size_t)k * n] = acc;
    }
    return;
  }
  uint64_t acc = 1;
  for (uint32_t k = 0; k < P.n_chunks; k++) {
    uint64_t np = 1, dp = 1;
    for (uint32_t j = k * P.chunk; j < (k + 1) * P.chunk && j < P.num_routed; j++) {
      uint64_t wv = gl::canon(P.wires[(size_t)j * n + i]);
      uint64_t num = fadd(fadd(wv, fmul(bx, P.k_is[j])), gamma);
      uint64_t den = fadd(fadd(wv, fmul(beta, gl::canon(P.sigmas[(size_t)j * n + i]))), gamma);
      np = fmul(np, num);
      dp = fmul(dp, den);
    }
    acc = fmul(acc, fmul(np, gl::inv(dp)));
    out[(size_t)k * n] = acc;
  }
}

// one CTA per challenge: exclusive prefix product of the row totals (last chunk column) -> z[ch][i]
__global__ void __launch_bounds__(1024) k_pp_scan(const uint64_t* __restrict__ local, uint32_t log_n, uint32_t n_chunks,
                                                   uint64_t* __restrict__ z) {
  __shared__ uint64_t part[1024];
  const size_t n = (size_t)1 << log_n;
  const uint32_t ch = blockIdx.x, tid = threadIdx.x;
  const uint64_t* q = local + ((size_t)ch * n_chunks + (n_chunks - 1)) * n;
  uint64_t* zo = z + (size_t)ch * n;
  const size_t per = (n + 1023) / 1024, lo = (size_t)tid * per, hi = lo + per < n ? lo + per : n;
  uint64_t p = 1;
  for (size_t i = lo; i < hi; i++) p = fmul(p, q[i]);
  part[tid] = p;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {  // inclusive scan of the partial products
    uint64_t v = tid >= off ? part[tid - off] : 1;
    __syncthreads();
    part[tid] = fmul(part[tid], v);
    __syncthreads();
  }
  uint64_t acc = tid ? part[tid - 1] : 1;
  for (size_t i = lo; i < hi; i++) {
    zo[i] = acc;
    acc = fmul(acc, q[i]);
  }
}

// out columns: [Z_0 .. Z_{c-1}, partial products of challenge 0 (n_chunks - 1), challenge 1, ...] x n
__global__ void __launch_bounds__(256) k_pp_finish(const uint64_t* __restrict__ local, const uint64_t* __restrict__ z,
                                                    uint32_t log_n, uint32_t n_chunks, uint32_t n_chal,
                                                    uint64_t* __restrict__ out) {
  const size_t n = (size_t)1 << log_n;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ch = blockIdx.y, npp = n_chunks - 1;
  const uint64_t zi = z[(size_t)ch * n + i];
  out[(size_t)ch * n + i] = zi;
  for (uint32_t k = 0; k < npp; k++)
    out[((size_t)n_chal + (size_t)ch * npp + k) * n + i] = fmul(zi, local[((size_t)ch * n_chunks + k) * n + i]);
}

// ------------------------------------------------------------------------------------------ gate evaluators
// Accumulates filter-free sum_q alpha^q c_q for every challenge — WITHOUT reducing: each product c_q * alpha^q is a
// 128-bit integer and a gate has at most a few hundred constraints, so the running sum fits five 32-bit words
// (sum < 2^136).  A push is 8 multiply-adds and 5 carry adds per challenge (a reduced multiply-add, gl::mad_nc, is
// 25 instructions), and the single reduction happens in value(): 2^128 = -2^32 (mod p).  Exact integer arithmetic,
// so the field element is the same as plonky2's reduce_with_powers.
struct Acc {
  uint32_t w[MAX_CHALLENGES][5];
  // powers of alpha (canonical) as the GATE table [term][MAX_CHALLENGES] (QuotientParams::apow_gates): one running
  // pointer, and the powers of all challenges for a term sit next to each other (one 16-byte load for two challenges)
  const uint64_t* ap;
  uint32_t n_chal;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int ch = 0; ch < MAX_CHALLENGES; ch++)
#pragma unroll
      for (int k = 0; k < 5; k++) w[ch][k] = 0;
  }
  __device__ __forceinline__ void add_product(int ch, uint64_t c, uint64_t a) {
    const uint32_t c0 = (uint32_t)c, c1 = (uint32_t)(c >> 32), a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %8, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %8, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %6, %7, %2;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "}"
        : "+r"(w[ch][0]), "+r"(w[ch][1]), "+r"(w[ch][2]), "+r"(w[ch][3]), "+r"(w[ch][4])
        : "r"(c0), "r"(c1), "r"(a0), "r"(a1));
  }
  __device__ __forceinline__ void weigh(const uint64_t* __restrict__ a, uint64_t c) {
    const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(a);
    add_product(0, c, a01.x);
    if (n_chal > 1) add_product(1, c, a01.y);
    if (n_chal > 2) {
      const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(a + 2);
      add_product(2, c, a23.x);
      if (n_chal > 3) add_product(3, c, a23.y);
    }
  }
  // c: any u64 (canonical or not); the alpha powers are canonical
  __device__ __forceinline__ void push(uint64_t c) {
    weigh(ap, c);
    ap += MAX_CHALLENGES;
  }
  // constraint number (current) + k of the gate, without advancing
  __device__ __forceinline__ void push_at(uint32_t k, uint64_t c) { weigh(ap + (size_t)k * MAX_CHALLENGES, c); }
  __device__ __forceinline__ void skip(uint32_t k) { ap += (size_t)k * MAX_CHALLENGES; }
  // the accumulated sum of challenge ch as a canonical field element
  __device__ __forceinline__ uint64_t value(int ch) const {
    // low 128 bits: x3:x2:x1:x0 -> (x1:x0) - x3 + x2 * (2^32 - 1), exactly gl::mul_nc's reduction
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 m,tl,th;\n\t"
        "sub.cc.u32 tl, %2, %5;\n\t"
        "subc.cc.u32 th, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 tl, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "mad.lo.cc.u32 tl, %4, 0xFFFFFFFF, tl;\n\t"
        "madc.hi.cc.u32 th, %4, 0xFFFFFFFF, th;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "add.u32 %1, th, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(w[ch][0]), "r"(w[ch][1]), "r"(w[ch][2]), "r"(w[ch][3]));
    // + w4 * 2^128 = - w4 * 2^32
    return gl::sub(gl::canon(gl::pack(r0, r1)), gl::pack(0u, w[ch][4]));
  }
};

struct Vars {
  const uint64_t* wires;  // column-major leaf-ordered LDE; element j of this point at wires[j * N]
  size_t N;
  const uint64_t* consts;  // ditto, already past the selector columns
  const uint64_t* pi_hash;
  ntt2::RootTables roots;
  __device__ __forceinline__ uint64_t w(uint32_t j) const { return wires[(size_t)j * N]; }
  __device__ __forceinline__ uint64_t c(uint32_t j) const { return consts[(size_t)j * N]; }
};

// prod_{x<4} (limb - x) = t (t + 2) with t = limb (limb - 3): two multiplications instead of three, the first a
// square.  The result is congruent to the product but NOT canonical — it only ever goes to Acc::push.
__device__ __forceinline__ uint64_t limb4_product(uint64_t limb) {
  const uint64_t t = gl::mul_nc(limb, fsub(limb, 3));  // limb canonical; t any u64 congruent to limb (limb - 3)
  return gl::mul_nc(t, gl::add_nc(t, 2));              // add_nc: one canonical operand is enough
}

// sum_j limb_j * 4^j (j < count <= 16) for canonical limbs, WITHOUT a reduction per term: the shifted limbs are added
// as a 128-bit integer (limb << 30 needs 94 bits, 16 of them 98) and reduced once.  7 instructions per limb instead of
// the 32 of comb = comb * 4 + limb in the field; the value is the same field element.
struct Base4Sum {
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  __device__ __forceinline__ void add(uint64_t limb, uint32_t j) {
    const uint32_t lo = (uint32_t)limb, hi = (uint32_t)(limb >> 32), sh = 2 * j;  // sh <= 30
    const uint32_t a0 = lo << sh, a1 = __funnelshift_l(lo, hi, sh), a2 = __funnelshift_l(hi, 0u, sh);
    asm("add.cc.u32 %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, %5;\n\t"
        "addc.cc.u32 %2, %2, %6;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3)
        : "r"(a0), "r"(a1), "r"(a2));
  }
  // canonical field element: (w1:w0) - w3 + w2 * (2^32 - 1)   [2^64 = 2^32 - 1, 2^96 = -1]
  __device__ __forceinline__ uint64_t value() const {
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 m,tl,th;\n\t"
        "sub.cc.u32 tl, %2, %5;\n\t"
        "subc.cc.u32 th, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 tl, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "mad.lo.cc.u32 tl, %4, 0xFFFFFFFF, tl;\n\t"
        "madc.hi.cc.u32 th, %4, 0xFFFFFFFF, th;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "add.u32 %1, th, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    return gl::canon(gl::pack(r0, r1));
  }
};

// sum_j x_j * 2^(s_j) for any u64 x_j and 0 <= s_j < 64, WITHOUT a reduction per term: the shifted terms (< 2^127) are
// added as a 160-bit integer (up to 2^32 terms fit) and reduced once, with Acc::value()'s reduction (2^128 = -2^32).
// Replaces the Horner forms sum = 2 sum + b (14 instructions per term) and sum = sum + coeff * b (33) of the bit / limb
// recombination constraints by 8 instructions per term; same field element (exact integer arithmetic).
// HI = false: shifts 0..31 (the term starts in word 0); HI = true: shifts 32..63 (word 1); `sh` = shift mod 32.
struct Sum160 {
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
  template <bool HI>
  __device__ __forceinline__ void add(uint64_t x, uint32_t sh) {
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    const uint32_t a0 = lo << sh, a1 = __funnelshift_l(lo, hi, sh), a2 = __funnelshift_l(hi, 0u, sh);
    if (!HI)
      asm("add.cc.u32 %0, %0, %5;\n\t"
          "addc.cc.u32 %1, %1, %6;\n\t"
          "addc.cc.u32 %2, %2, %7;\n\t"
          "addc.cc.u32 %3, %3, 0;\n\t"
          "addc.u32 %4, %4, 0;"
          : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "+r"(w4)
          : "r"(a0), "r"(a1), "r"(a2));
    else
      asm("add.cc.u32 %0, %0, %4;\n\t"
          "addc.cc.u32 %1, %1, %5;\n\t"
          "addc.cc.u32 %2, %2, %6;\n\t"
          "addc.u32 %3, %3, 0;"
          : "+r"(w1), "+r"(w2), "+r"(w3), "+r"(w4)
          : "r"(a0), "r"(a1), "r"(a2));
  }
  // shift s in 0..63, known to lie on one side of 32 for the whole loop that calls it
  __device__ __forceinline__ void add_shift(uint64_t x, uint32_t s) {
    if (s < 32)
      add<false>(x, s);
    else
      add<true>(x, s - 32);
  }
  __device__ __forceinline__ uint64_t value() const {
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 m,tl,th;\n\t"
        "sub.cc.u32 tl, %2, %5;\n\t"
        "subc.cc.u32 th, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 tl, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "mad.lo.cc.u32 tl, %4, 0xFFFFFFFF, tl;\n\t"
        "madc.hi.cc.u32 th, %4, 0xFFFFFFFF, th;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, tl, m;\n\t"
        "subc.u32 th, th, 0;\n\t"
        "add.u32 %1, th, m;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
    return gl::sub(gl::canon(gl::pack(r0, r1)), gl::pack(0u, w4));  // + w4 * 2^128 = - w4 * 2^32
  }
};

__device__ __forceinline__ void mds_plain(uint64_t (&s)[12]) {
  poseidon::mds_layer(s, poseidon::RCF + 24 * 29);  // round-30 "constants" are zero: pure MDS
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::canon(s[i]);
}
__device__ __forceinline__ uint64_t sbox7c(uint64_t x) { return gl::canon(poseidon::sbox7(x)); }

// PoseidonGate (gates/poseidon.rs eval_unfiltered): the permutation with every S-box input replaced by a wire and the
// difference to the computed input as a constraint.  plonky2 evaluates it in the "fast partial round" form; the S-box
// inputs of that form are those of the plain permutation (lane 0 after the constant layer: the change of basis of the
// fast form touches lanes 1..11 only and commutes with anything done to lane 0 — checked for the witness generator by
// tests/test_plonk_oracle.py and, value for value on random points, by the GPU parity tests against the oracle's
// fast-form evaluator), so the evaluator IS the leaf-hash permutation (poseidon.cuh, v6 schedule: MDS layers on the
// FP64 pipe, two partial rounds per step) with hooks: ~21k instead of ~38k instructions per point with the fast-form
// tables and 64-bit constant multiplications.
__device__ void eval_poseidon_gate(const Vars& v, Acc& acc) {
  constexpr int SWAP = 24, DELTA = 25, FULL0 = 29, PARTIAL = 65, FULL1 = 87;
  const uint64_t swap = v.w(SWAP);
  acc.push(fmul(swap, fsub(swap, 1)));
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t l = v.w(i), r = v.w(i + 4), d = v.w(DELTA + i);
    acc.push(fsub(fmul(swap, fsub(r, l)), d));
    s[i] = fadd(l, d);
    s[i + 4] = fsub(r, d);
  }
#pragma unroll
  for (int i = 8; i < 12; i++) s[i] = v.w(i);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_nc(s[i], poseidon::RC[i]);
  // first full rounds: s = S-box inputs (round constants included by the previous layer); wires from round 1 on
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    if (r != 0) {
#pragma unroll
      for (int i = 0; i < 12; i++) {
        const uint64_t in = v.w(FULL0 + 12 * (r - 1) + i);
        acc.push(gl::sub_nc(s[i], in));  // any u64 minus a canonical wire: congruent, which is all push needs
        s[i] = in;
      }
    }
    poseidon::full_round_v6(s, poseidon::RC6 + 24 * r);
  }
  {
    double zlo[12], zhi[12];
    zlo[0] = zhi[0] = 0.;
#pragma unroll
    for (int i = 1; i < 12; i++) poseidon::limbs_from_u64(s[i], zlo[i], zhi[i]);
    uint64_t s0 = s[0];
    const uint32_t vz = poseidon::lane_varying_zero();
#pragma unroll 1
    for (int r = 0; r < 22; r += 2) {
      const uint64_t in0 = v.w(PARTIAL + r), in1 = v.w(PARTIAL + r + 1);
      acc.push(gl::sub_nc(s0, in0));
      s0 = in0;
      uint64_t mid_computed = 0;
      poseidon::P2B_PAIR_HOOK(s0, zlo, zhi, 4 + r, vz, [&](uint64_t x) {
        mid_computed = x;
        return in1;
      });
      acc.push(gl::sub_nc(mid_computed, in1));
    }
    s[0] = s0;
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = gl::sub_nc(poseidon::fold_f64(zlo[i], zhi[i]), (uint64_t)P2B_LAZY_OFFSET);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const uint64_t in = v.w(FULL1 + 12 * r + i);
      acc.push(gl::sub_nc(s[i], in));
      s[i] = in;
    }
    poseidon::full_round_v6(s, poseidon::RC6 + 24 * (26 + r));
  }
#pragma unroll
  for (int i = 0; i < 12; i++) acc.push(gl::sub_nc(s[i], v.w(12 + i)));
}

__device__ __forceinline__ void eval_gate_light(const Gate& g, const Vars& v, Acc& acc) {
  switch (g.kind) {
    case GATE_NOOP:
      break;
    case GATE_CONSTANT:
      for (uint32_t i = 0; i < g.p0; i++) acc.push(fsub(v.c(i), v.w(i)));
      break;
    case GATE_PUBLIC_INPUT:
      for (uint32_t i = 0; i < 4; i++) acc.push(fsub(v.w(i), v.pi_hash[i]));
      break;
    case GATE_ARITHMETIC: {
      const uint64_t c0 = v.c(0), c1 = v.c(1);
      for (uint32_t i = 0; i < g.p0; i++) {
        uint64_t computed = fadd(fmul(fmul(v.w(4 * i), v.w(4 * i + 1)), c0), fmul(v.w(4 * i + 2), c1));
        acc.push(fsub(v.w(4 * i + 3), computed));
      }
      break;
    }
    case GATE_BASE_SUM: {
      if (g.p0 <= 64) {  // sum_i limb_i 2^i, one reduction
        Sum160 sum;
        const uint32_t lo_n = g.p0 < 32 ? g.p0 : 32;
        for (uint32_t i = 0; i < lo_n; i++) sum.add<false>(v.w(1 + i), i);
        for (uint32_t i = 32; i < g.p0; i++) sum.add<true>(v.w(1 + i), i - 32);
        acc.push(fsub(sum.value(), v.w(0)));
      } else {
        uint64_t sum = 0;
        for (uint32_t i = g.p0; i-- > 0;) sum = fadd(fadd(sum, sum), v.w(1 + i));
        acc.push(fsub(sum, v.w(0)));
      }
      for (uint32_t i = 0; i < g.p0; i++) {
        uint64_t l = v.w(1 + i);
        acc.push(gl::mul_nc(l, fsub(l, 1)));  // push takes any u64
      }
      break;
    }
    case GATE_U32_ARITHMETIC: {
      const uint32_t ops = g.p0;
#pragma unroll 1
      for (uint32_t i = 0; i < ops; i++) {
        uint64_t lo = v.w(6 * i + 3), hi = v.w(6 * i + 4), inv = v.w(6 * i + 5);
        uint64_t computed = fadd(fmul(v.w(6 * i), v.w(6 * i + 1)), v.w(6 * i + 2));
        acc.push(fmul(fsub(fmul(inv, fsub(0xFFFFFFFFull, hi)), 1), lo));
        acc.push(fsub(fadd(fmul(hi, 1ull << 32), lo), computed));
        // range checks in the reference's order (limb 31 first: each limb's product is constraint 2 + (31 - j))
        Base4Sum clo, chi;
#pragma unroll 2
        for (int j = 31; j >= 0; j--) {
          const uint64_t limb = v.w(6 * ops + 32 * i + j);
          acc.push(limb4_product(limb));
          if (j < 16)
            clo.add(limb, (uint32_t)j);
          else
            chi.add(limb, (uint32_t)(j - 16));
        }
        acc.push(fsub(clo.value(), lo));
        acc.push(fsub(chi.value(), hi));
      }
      break;
    }
    case GATE_U32_ADD_MANY: {
      const uint32_t na = g.p0, ops = g.p1, per = na + 3;
#pragma unroll 1
      for (uint32_t i = 0; i < ops; i++) {
        uint64_t computed = 0;
        for (uint32_t j = 0; j <= na; j++) computed = fadd(computed, v.w(per * i + j));  // addends + carry
        uint64_t res = v.w(per * i + na + 1), carry = v.w(per * i + na + 2);
        acc.push(fsub(fadd(fmul(carry, 1ull << 32), res), computed));
        Base4Sum cres, ccar;
#pragma unroll 2
        for (int j = 17; j >= 0; j--) {
          const uint64_t limb = v.w(per * ops + 18 * i + j);
          acc.push(limb4_product(limb));
          if (j < 16)
            cres.add(limb, (uint32_t)j);
          else
            ccar.add(limb, (uint32_t)(j - 16));
        }
        acc.push(fsub(cres.value(), res));
        acc.push(fsub(ccar.value(), carry));
      }
      break;
    }
    case GATE_U32_SUBTRACTION: {
      const uint32_t ops = g.p0;
#pragma unroll 1
      for (uint32_t i = 0; i < ops; i++) {
        uint64_t res = v.w(5 * i + 3), bout = v.w(5 * i + 4);
        uint64_t initial = fsub(fsub(v.w(5 * i), v.w(5 * i + 1)), v.w(5 * i + 2));
        acc.push(fsub(res, fadd(initial, fmul(bout, 1ull << 32))));
        Base4Sum comb;
#pragma unroll 2
        for (int j = 15; j >= 0; j--) {
          const uint64_t limb = v.w(5 * ops + 16 * i + j);
          acc.push(limb4_product(limb));
          comb.add(limb, (uint32_t)j);
        }
        acc.push(fsub(comb.value(), res));
        acc.push(fmul(bout, fsub(1, bout)));
      }
      break;
    }
    case GATE_U32_RANGE_CHECK: {
      const uint32_t nl = g.p0;
#pragma unroll 1
      for (uint32_t i = 0; i < nl; i++) {
        Base4Sum comb;
#pragma unroll 4
        for (int j = 15; j >= 0; j--) comb.add(v.w(nl + 16 * i + j), (uint32_t)j);
        acc.push(fsub(comb.value(), v.w(i)));
#pragma unroll 2
        for (int j = 0; j < 16; j++) acc.push(limb4_product(v.w(nl + 16 * i + j)));
      }
      break;
    }
    case GATE_U32_INTERLEAVE: {  // 32 big-endian bits per op after the 2 * ops routed wires
      const uint32_t ops = g.p0;
#pragma unroll 1
      for (uint32_t i = 0; i < ops; i++) {
        // big-endian bits: x = sum_j b_j 2^(31 - j), x_interleaved = sum_j b_j 4^(31 - j)
        Sum160 cx, cxi;
#pragma unroll 2
        for (int j = 0; j < 16; j++) {
          const uint64_t b = v.w(2 * ops + 32 * i + j);
          cx.add<false>(b, (uint32_t)(31 - j));
          cxi.add<true>(b, (uint32_t)(30 - 2 * j));  // 2 (31 - j) - 32
        }
#pragma unroll 2
        for (int j = 16; j < 32; j++) {
          const uint64_t b = v.w(2 * ops + 32 * i + j);
          cx.add<false>(b, (uint32_t)(31 - j));
          cxi.add<false>(b, (uint32_t)(62 - 2 * j));
        }
        acc.push(fsub(cx.value(), v.w(2 * i)));
        acc.push(fsub(cxi.value(), v.w(2 * i + 1)));
#pragma unroll 2
        for (int j = 0; j < 32; j++) {
          const uint64_t b = v.w(2 * ops + 32 * i + j);
          acc.push(gl::mul_nc(b, fsub(b, 1)));
        }
      }
      break;
    }
    case GATE_UNINTERLEAVE_TO_U32:
    case GATE_UNINTERLEAVE_TO_B32: {  // 64 big-endian bits per op after the 3 * ops routed wires
      const uint32_t ops = g.p0;
      const bool b32 = g.kind == GATE_UNINTERLEAVE_TO_B32;
#pragma unroll 1
      for (uint32_t i = 0; i < ops; i++) {
        // x = sum_j b_j 2^(63 - j); evens / odds = sum_j b_{2j}, b_{2j+1} times 2^(31 - j) (U32) or 4^(31 - j) (B32)
        Sum160 cx, ce, co;
#pragma unroll 4
        for (int j = 0; j < 32; j++) cx.add<true>(v.w(3 * ops + 64 * i + j), (uint32_t)(31 - j));
#pragma unroll 4
        for (int j = 32; j < 64; j++) cx.add<false>(v.w(3 * ops + 64 * i + j), (uint32_t)(63 - j));
        acc.push(fsub(cx.value(), v.w(3 * i)));
        if (b32) {
#pragma unroll 2
          for (int j = 0; j < 16; j++) {
            ce.add<true>(v.w(3 * ops + 64 * i + 2 * j), (uint32_t)(30 - 2 * j));
            co.add<true>(v.w(3 * ops + 64 * i + 2 * j + 1), (uint32_t)(30 - 2 * j));
          }
#pragma unroll 2
          for (int j = 16; j < 32; j++) {
            ce.add<false>(v.w(3 * ops + 64 * i + 2 * j), (uint32_t)(62 - 2 * j));
            co.add<false>(v.w(3 * ops + 64 * i + 2 * j + 1), (uint32_t)(62 - 2 * j));
          }
        } else {
#pragma unroll 2
          for (int j = 0; j < 32; j++) {
            ce.add<false>(v.w(3 * ops + 64 * i + 2 * j), (uint32_t)(31 - j));
            co.add<false>(v.w(3 * ops + 64 * i + 2 * j + 1), (uint32_t)(31 - j));
          }
        }
        acc.push(fsub(ce.value(), v.w(3 * i + 1)));
        acc.push(fsub(co.value(), v.w(3 * i + 2)));
#pragma unroll 2
        for (int j = 0; j < 64; j++) {
          const uint64_t b = v.w(3 * ops + 64 * i + j);
          acc.push(gl::mul_nc(b, fsub(b, 1)));
        }
      }
      break;
    }
    case GATE_COMPARISON: {  // p0 = num_bits, p1 = num_chunks
      const uint32_t nc = g.p1, cb = (g.p0 + nc - 1) / nc;
      uint64_t f_comb = 0, s_comb = 0;
      for (uint32_t i = nc; i-- > 0;) {
        f_comb = fadd(fmul(f_comb, 1ull << cb), v.w(4 + i));
        s_comb = fadd(fmul(s_comb, 1ull << cb), v.w(4 + nc + i));
      }
      acc.push(fsub(f_comb, v.w(0)));
      acc.push(fsub(s_comb, v.w(1)));
      uint64_t msd = 0;
      for (uint32_t i = 0; i < nc; i++) {
        const uint64_t fc = v.w(4 + i), sc = v.w(4 + nc + i);
        if (cb == 2) {  // prod_{x<4} (chunk - x): the range-check product of the u32 gates
          acc.push(limb4_product(fc));
          acc.push(limb4_product(sc));
        } else {
          uint64_t fp = 1, sp = 1;
          for (uint64_t x = 0; x < (1ull << cb); x++) {
            fp = fmul(fp, fsub(fc, x));
            sp = fmul(sp, fsub(sc, x));
          }
          acc.push(fp);
          acc.push(sp);
        }
        const uint64_t diff = fsub(sc, fc);
        const uint64_t dummy = v.w(4 + 2 * nc + i), eq = v.w(4 + 3 * nc + i), inter = v.w(4 + 4 * nc + i);
        acc.push(fsub(fmul(diff, dummy), fsub(1, eq)));
        acc.push(gl::mul_nc(eq, diff));
        acc.push(fsub(inter, fmul(eq, msd)));
        msd = fadd(inter, fmul(fsub(1, eq), diff));
      }
      const uint64_t msd_w = v.w(3);
      acc.push(fsub(msd_w, msd));
      uint64_t comb = 0;
      for (uint32_t b = 0; b <= cb; b++) {
        const uint64_t bit = v.w(4 + 5 * nc + b);
        acc.push(gl::mul_nc(bit, fsub(1, bit)));
      }
      for (uint32_t b = cb + 1; b-- > 0;) comb = fadd(fadd(comb, comb), v.w(4 + 5 * nc + b));
      acc.push(fsub(fadd(1ull << cb, msd_w), comb));
      acc.push(fsub(v.w(2), v.w(4 + 5 * nc + cb)));
      break;
    }
    case GATE_ARITHMETIC_EXT:
    case GATE_MUL_EXT: {  // extension elements in wire pairs; output - (m0 m1 c0 [+ addend c1]) per component
      const bool arith = g.kind == GATE_ARITHMETIC_EXT;
      const uint32_t per = arith ? 8 : 6;
      const uint64_t c0 = v.c(0), c1 = arith ? v.c(1) : 0;
      for (uint32_t i = 0; i < g.p0; i++) {
        const uint32_t q = per * i;
        const gl::ext2 prod = gl::ext_mul(gl::ext2{v.w(q), v.w(q + 1)}, gl::ext2{v.w(q + 2), v.w(q + 3)});
        uint64_t r0 = fmul(prod.c0, c0), r1 = fmul(prod.c1, c0);
        if (arith) r0 = fadd(r0, fmul(v.w(q + 4), c1)), r1 = fadd(r1, fmul(v.w(q + 5), c1));
        acc.push(fsub(v.w(q + per - 2), r0));
        acc.push(fsub(v.w(q + per - 1), r1));
      }
      break;
    }
    case GATE_REDUCING:
    case GATE_REDUCING_EXT: {  // output 0..2, alpha 2..4, old_acc 4..6, coefficients from 6, accumulators after them
      const uint32_t n = g.p0;
      const bool ext = g.kind == GATE_REDUCING_EXT;
      const uint32_t start_accs = 6 + (ext ? 2 * n : n);
      const gl::ext2 alpha{v.w(2), v.w(3)};
      const uint64_t alpha1_7 = fmul(7, alpha.c1);  // acc * alpha with 7 alpha_1 formed once: two dot products per step
      gl::ext2 a{v.w(4), v.w(5)};
      for (uint32_t i = 0; i < n; i++) {
        const gl::ext2 t{gl::dot2(a.c0, alpha.c0, a.c1, alpha1_7), gl::dot2(a.c0, alpha.c1, a.c1, alpha.c0)};
        const uint32_t nx = i == n - 1 ? 0 : start_accs + 2 * i;
        const gl::ext2 nxt{v.w(nx), v.w(nx + 1)};
        const uint64_t k0 = ext ? v.w(6 + 2 * i) : v.w(6 + i), k1 = ext ? v.w(7 + 2 * i) : 0;
        acc.push(fsub(fadd(t.c0, k0), nxt.c0));
        acc.push(fsub(fadd(t.c1, k1), nxt.c1));
        a = nxt;
      }
      break;
    }
    default:
      break;
  }
}

__device__ __forceinline__ void eval_gate_heavy(const Gate& g, const Vars& v, Acc& acc) {
  switch (g.kind) {
    case GATE_POSEIDON:
      eval_poseidon_gate(v, acc);
      break;
    case GATE_RANDOM_ACCESS: {  // p0 = bits (<= 6), p1 = num_copies | num_extra_constants << 16
      const uint32_t bits = g.p0, copies = g.p1 & 0xFFFF, extra = g.p1 >> 16, vec = 1u << bits;
      const uint32_t routed = (2 + vec) * copies + extra;
      for (uint32_t cp = 0; cp < copies; cp++) {
        const uint32_t q = (2 + vec) * cp, bq = routed + cp * bits;
        uint64_t rec = 0, items[64];
        for (uint32_t t = 0; t < bits; t++) {
          const uint64_t b = v.w(bq + t);
          acc.push(fmul(b, fsub(b, 1)));
        }
        for (uint32_t t = bits; t-- > 0;) rec = fadd(fadd(rec, rec), v.w(bq + t));
        acc.push(fsub(rec, v.w(q)));
        for (uint32_t t = 0; t < vec; t++) items[t] = v.w(q + 2 + t);
        uint32_t len = vec;
        for (uint32_t t = 0; t < bits; t++) {
          const uint64_t b = v.w(bq + t);
          len >>= 1;
          for (uint32_t u = 0; u < len; u++) items[u] = fadd(items[2 * u], fmul(b, fsub(items[2 * u + 1], items[2 * u])));
        }
        acc.push(fsub(items[0], v.w(q + 1)));
      }
      for (uint32_t t = 0; t < extra; t++) acc.push(fsub(v.c(t), v.w((2 + vec) * copies + t)));
      break;
    }
    case GATE_POSEIDON_MDS: {  // the MDS layer on 12 extension elements, componentwise
      for (int t = 0; t < 2; t++) {
        uint64_t s[12];
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = v.w(2 * i + t);
        mds_plain(s);
        // constraints are ordered (row, component): remember the accumulator slot of each
        for (int r = 0; r < 12; r++) {
          const uint64_t cst = fsub(s[r], v.w(24 + 2 * r + t));
          acc.push_at(2 * r + t, cst);
        }
      }
      acc.skip(24);
      break;
    }
    case GATE_COSET_INTERPOLATION: {  // p0 = subgroup_bits, p1 = degree; barycentric weights of the subgroup = x_i / n
      const uint32_t n = 1u << g.p0, degree = g.p1, n_int = (n - 2) / (degree - 1);
      const uint32_t pt = 1 + 2 * n, val = pt + 2, ie0 = pt + 4, ip0 = ie0 + 2 * n_int, sh = ie0 + 4 * n_int;
      const uint64_t gen = ntt2::root_pow(v.roots, g.p0, 1), ninv = GL_P - ((GL_P - 1) >> g.p0);
      const uint64_t shift = v.w(0);
      const gl::ext2 z{v.w(sh), v.w(sh + 1)};
      acc.push(fsub(v.w(pt), fmul(z.c0, shift)));
      acc.push(fsub(v.w(pt + 1), fmul(z.c1, shift)));
      gl::ext2 ev{0, 0}, pr{1, 0};
      uint64_t x = 1;
      uint32_t i = 0, hi = degree;
      for (uint32_t c = 0; c <= n_int; c++) {
        for (; i < hi && i < n; i++) {  // partial_interpolate
          const gl::ext2 term{fsub(z.c0, x), z.c1};
          const uint64_t wt = fmul(x, ninv);
          const gl::ext2 wv{fmul(v.w(1 + 2 * i), wt), fmul(v.w(2 + 2 * i), wt)};
          ev = gl::ext_add(gl::ext_mul(ev, term), gl::ext_mul(wv, pr));
          pr = gl::ext_mul(pr, term);
          x = fmul(x, gen);
        }
        if (c == n_int) break;
        const gl::ext2 ie{v.w(ie0 + 2 * c), v.w(ie0 + 2 * c + 1)}, ip{v.w(ip0 + 2 * c), v.w(ip0 + 2 * c + 1)};
        acc.push(fsub(ie.c0, ev.c0));
        acc.push(fsub(ie.c1, ev.c1));
        acc.push(fsub(ip.c0, pr.c0));
        acc.push(fsub(ip.c1, pr.c1));
        ev = ie;
        pr = ip;
        hi = 1 + (degree - 1) * (c + 1) + degree - 1;  // the next chunk starts where this one ended
      }
      acc.push(fsub(v.w(val), ev.c0));
      acc.push(fsub(v.w(val + 1), ev.c1));
      break;
    }
    default:
      break;
  }
}


// ------------------------------------------------------------------------------------------ quotient values
struct QuotientParams {
  const uint64_t* cs_lde;     // constants || sigmas, column-major, leaf order, column stride N
  const uint64_t* wires_lde;
  const uint64_t* zs_lde;     // Zs || partial products
  size_t N;                   // n << rate_bits
  const Gate* gates;
  const uint64_t* k_is;
  const uint64_t* apow;       // [challenge][n_terms] powers of alpha
  const uint64_t* apow_gates; // [num_gate_constraints][MAX_CHALLENGES]: alpha_c^(n_chal * (num_pp + 2) + q), the gate terms
  const uint64_t* zh;         // [2^mdb] Z_H on the coset, then [2^mdb] inverses
  uint64_t* parts;            // [1 + n_gates][challenge][lde_size], LEAF order (k_quotient_combine un-reverses)
  const uint64_t* betas;      // device
  const uint64_t* gammas;     // device
  const uint64_t* pi_hash;    // device, 4 canonical elements
  uint32_t degree_bits, mdb;  // lde_size = 2^(degree_bits + mdb)
  uint32_t num_routed, num_constants, num_selectors, n_chal, chunk, num_pp, n_gates, n_terms;
  uint32_t point_major, list_len;  // k_quotient_gates' CTA -> (point block, gate) mapping, see there
  // EXT launches of k_quotient_gates (low-degree gates, see there): the unfiltered sums of the first 2^ext_log_pts
  // leaves go to ext_out[list position][challenge][2^ext_log_pts] in NATURAL order of that sub-coset
  uint64_t* ext_out;
  uint32_t ext_log_pts;
  ntt2::RootTables roots;
};

// One thread = (point, term group): group 0 is the permutation argument (L_0 (Z - 1) and the partial-product checks),
// group 1 + g is gate g.  A point's work is a long dependent instruction stream (the PoseidonGate alone is a whole
// permutation), and one thread per point leaves a 2^12-row proof with 7 warps per SM; splitting by term group puts
// (1 + n_gates) times as many independent streams in flight.  Each group writes its alpha-weighted sum to
// parts[group][challenge][leaf]; k_quotient_combine adds them (field addition is exact, so the order is irrelevant)
// and divides by Z_H.
// The groups are evaluated by separate kernels, so that each gets the register budget its code needs (one kernel for
// everything had to live in 64 registers and spilled ~2 KB per thread in the widest gates):
//   k_quotient_perm                 grid (lde_size / 128, 1)          the permutation argument
//   k_quotient_gates<false, false>  grid (lde_size / 128, n_light)    light gates evaluated at every point
//   k_quotient_gates<true, false>   grid (lde_size / 128, n_heavy)    Poseidon, PoseidonMds, RandomAccess, CosetInterpolation
//                                                                     (state arrays: 12-element permutation state, 64 items)
//   k_quotient_gates<false, true>   grid (D n / 128, n_ext[D])        gates of constraint degree <= D in {2, 4}: the first D n
//                                                                     leaves only, unfiltered, extended by NTT (see there)
// gate_list[blockIdx.y] = index of the gate in P.gates.
__device__ __forceinline__ bool gate_is_heavy(uint32_t kind) {
  return kind == GATE_POSEIDON || kind == GATE_POSEIDON_MDS || kind == GATE_RANDOM_ACCESS || kind == GATE_COSET_INTERPOLATION;
}
#ifndef P2B_QUOT_MINB
#define P2B_QUOT_MINB 6  // 80 registers: no spills in the light gates (8 -> 64 registers spilled ~300 B); measured +1.6 % proofs/s
#endif
#ifndef P2B_QUOT_HEAVY_MINB
#define P2B_QUOT_HEAVY_MINB 4
#endif

// compute_filter (gates/gate.rs): prod_{q in the gate's selector group, q != row} (q - s) * (UNUSED_SELECTOR - s)
__device__ __forceinline__ uint64_t gate_filter(const Gate& gate, uint64_t s, uint32_t num_selectors) {
  uint64_t filter = 1;
  for (uint32_t q = gate.group_start; q < gate.group_end; q++)
    if (q != gate.row) filter = fmul(filter, fsub(q, s));
  if (num_selectors > 1) filter = fmul(filter, fsub(P2B_UNUSED_SELECTOR, s));
  return filter;
}

__global__ void __launch_bounds__(128, 6) k_quotient_perm(QuotientParams P) {
  const uint32_t log_lde = P.degree_bits + P.mdb;
  const size_t lde_size = (size_t)1 << log_lde;
  const size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= lde_size) return;
  const size_t i = ntt2::brev((uint32_t)leaf, log_lde);
  const uint32_t nch = P.n_chal, npp = P.num_pp;
  const uint64_t* cs = P.cs_lde + leaf;
  const uint64_t* wr = P.wires_lde + leaf;
  const size_t N = P.N;
  uint64_t res[MAX_CHALLENGES];
#pragma unroll
  for (int c = 0; c < MAX_CHALLENGES; c++) res[c] = 0;
  {
    const size_t i_next = (i + ((size_t)1 << P.mdb)) & (lde_size - 1);
    const size_t leaf_next = ntt2::brev((uint32_t)i_next, log_lde);
    const uint32_t zi = (uint32_t)(i & (((size_t)1 << P.mdb) - 1));
    const uint64_t x = fmul(7, ntt2::root_pow(P.roots, log_lde, i));
    const uint64_t z_h = P.zh[zi];
    const uint64_t l0 = fmul(z_h, gl::inv(fmul((uint64_t)1 << P.degree_bits, fsub(x, 1))));
    const uint64_t* zs = P.zs_lde + leaf;
    // vanishing_z_1_terms (terms 0 .. nch-1), then the partial-product checks of every challenge
    // (terms nch + cc * (npp + 1) + k); each term is weighted by every challenge's own power of alpha
    for (uint32_t k = 0; k < nch; k++) {
      const uint64_t term = fmul(l0, fsub(gl::canon(zs[(size_t)k * N]), 1));
      for (uint32_t c = 0; c < nch; c++) res[c] = gl::mad_nc(P.apow[(size_t)c * P.n_terms + k], term, res[c]);
    }
    // Every routed wire and its sigma are loaded ONCE for all challenges, four wires at a time before any of them is
    // used (the loads of this loop were the largest long-scoreboard stall of the kernel: one dependent DRAM round
    // trip per wire and challenge), and the challenges' products are independent multiply chains.
    uint64_t beta[MAX_CHALLENGES], gamma[MAX_CHALLENGES], bx[MAX_CHALLENGES];
#pragma unroll
    for (int cc = 0; cc < MAX_CHALLENGES; cc++)
      if (cc < (int)nch) {
        beta[cc] = gl::canon(P.betas[cc]), gamma[cc] = gl::canon(P.gammas[cc]);
        bx[cc] = fmul(beta[cc], x);
      }
    for (uint32_t k = 0; k <= npp; k++) {
      uint64_t np[MAX_CHALLENGES], dp[MAX_CHALLENGES];
#pragma unroll
      for (int cc = 0; cc < MAX_CHALLENGES; cc++) np[cc] = 1, dp[cc] = 1;
      const uint32_t j_end = (k + 1) * P.chunk < P.num_routed ? (k + 1) * P.chunk : P.num_routed;
      for (uint32_t j0 = k * P.chunk; j0 < j_end; j0 += 4) {
        uint64_t wv[4], sg[4], kj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t j = j0 + u < j_end ? j0 + u : j_end - 1;  // clamped: the tail re-reads the last wire, unused
          wv[u] = wr[(size_t)j * N];
          sg[u] = cs[(size_t)(P.num_constants + j) * N];
          kj[u] = P.k_is[j];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (j0 + u >= j_end) break;
          const uint64_t w_c = gl::canon(wv[u]), s_c = gl::canon(sg[u]);
          // w + gamma is shared by the numerator and the denominator factor; beta k x + (w + gamma) is one
          // multiply-add with a single reduction, and the running products stay non-canonical u64 (mul_nc takes
          // any u64): 101 instead of 132 instructions per wire and challenge
#pragma unroll
          for (int cc = 0; cc < MAX_CHALLENGES; cc++)
            if (cc < (int)nch) {
              const uint64_t wg = fadd(w_c, gamma[cc]);
              np[cc] = gl::mul_nc(np[cc], gl::mad_nc(bx[cc], kj[u], wg));
              dp[cc] = gl::mul_nc(dp[cc], gl::mad_nc(beta[cc], s_c, wg));
            }
        }
      }
#pragma unroll
      for (int cc = 0; cc < MAX_CHALLENGES; cc++)
        if (cc < (int)nch) {
          const uint64_t prev = k == 0 ? zs[(size_t)cc * N] : zs[(size_t)(nch + cc * npp + k - 1) * N];
          const uint64_t next = k == npp ? P.zs_lde[(size_t)cc * N + leaf_next] : zs[(size_t)(nch + cc * npp + k) * N];
          const uint64_t term = fsub(fmul(gl::canon(prev), np[cc]), fmul(gl::canon(next), dp[cc]));
          const uint32_t idx = nch + cc * (npp + 1) + k;
          for (uint32_t c = 0; c < nch; c++) res[c] = gl::mad_nc(P.apow[(size_t)c * P.n_terms + idx], term, res[c]);
        }
    }
    for (uint32_t c = 0; c < nch; c++) res[c] = gl::canon(res[c]);  // k_quotient_combine adds canonical parts
  }
  for (uint32_t c = 0; c < nch; c++) P.parts[(size_t)c * lde_size + leaf] = res[c];
}

// EXT launches — low-degree gates.  The alpha-weighted constraint sum G_g(x) of gate g, WITHOUT its selector filter, is
// a polynomial of degree <= d_g (n - 1) in x, where d_g is the gate's constraint degree (wires and constants are
// polynomials of degree < n).  For d_g <= D < 2^mdb it is therefore fixed by its values on the sub-coset 7 <w_{D n}> —
// the FIRST D n leaves of the leaf-ordered LDE — and its values on the whole quotient coset follow by one inverse NTT
// of size D n and one forward NTT of size 2^mdb n (both cosets have the shift 7, so no rescaling); the filter, a
// polynomial in the selector column, is multiplied in by k_quotient_combine at all points.  Field arithmetic is exact,
// so the extended values ARE the directly evaluated ones (same field elements, bit for bit): a degree-2 gate
// (booleans, base sums, reducing) costs a quarter and a degree-3/4 gate (arithmetic, u32 limbs, comparison) half of
// the evaluator work, for two NTT columns per gate.
template <bool HEAVY, bool EXT>
__global__ void __launch_bounds__(128, HEAVY ? P2B_QUOT_HEAVY_MINB : P2B_QUOT_MINB)
k_quotient_gates(QuotientParams P, const uint32_t* __restrict__ gate_list) {
  const uint32_t log_lde = P.degree_bits + P.mdb;
  const size_t lde_size = (size_t)1 << log_lde;
  const size_t n_pts = EXT ? (size_t)1 << P.ext_log_pts : lde_size;
  // gate-major (grid (points / 128, gates), the default): the CTAs resident together evaluate the same few gates — the
  // smallest instruction footprint.  point-major (grid (gates * points / 128, 1), gate fastest; P2B_QUOT_POINT_MAJOR=1):
  // the CTAs resident together cover the same points for ALL gates, so a large circuit's wires are read from HBM once per
  // kernel instead of once per gate (2^16 rows: 4.7 GB -> ~1 GB) — measured slower at every size (p2b.cu quotient_core).
  const uint32_t list_pos = P.point_major ? blockIdx.x % P.list_len : blockIdx.y;
  const size_t point_block = P.point_major ? blockIdx.x / P.list_len : blockIdx.x;
  const size_t leaf = point_block * blockDim.x + threadIdx.x;
  if (leaf >= n_pts) return;
  const uint32_t nch = P.n_chal;
  const uint64_t* cs = P.cs_lde + leaf;
  const size_t N = P.N;
  const uint32_t g = gate_list[list_pos];
  // gate constraints: every gate's constraint q lands on term nch*(npp+2) + q
  Vars v{P.wires_lde + leaf, N, cs + (size_t)P.num_selectors * N, P.pi_hash, P.roots};
  const Gate gate = P.gates[g];
  Acc acc;
  acc.clear();
  acc.ap = P.apow_gates;
  acc.n_chal = nch;
  if (HEAVY)
    eval_gate_heavy(gate, v, acc);
  else
    eval_gate_light(gate, v, acc);
  if (EXT) {
    const size_t idx = ntt2::brev((uint32_t)leaf, P.ext_log_pts);
#pragma unroll
    for (int c = 0; c < MAX_CHALLENGES; c++)
      if (c < (int)nch) P.ext_out[((size_t)list_pos * nch + c) * n_pts + idx] = acc.value(c);
    return;
  }
  const uint64_t filter = gate_filter(gate, gl::canon(cs[(size_t)gate.selector_index * N]), P.num_selectors);
#pragma unroll
  for (int c = 0; c < MAX_CHALLENGES; c++)
    if (c < (int)nch) P.parts[((size_t)(1 + g) * nch + c) * lde_size + leaf] = fmul(filter, acc.value(c));
}

// out[c][i] = (sum over the direct parts + sum over the extended gates of filter * extended sum) / Z_H(x_i); one thread
// per LEAF (parts, extended sums and selector columns are leaf-ordered: coalesced), the quotient values are stored in
// natural order (i = bit-reversed leaf) for the coset iNTT that follows.
//   part_list[0 .. n_direct)          indices into parts (0 = the permutation argument, 1 + g = gate g)
//   ext_list[0 .. n_ext)              gate index of the extended sum ext[k] ([n_ext][challenge][lde_size], leaf order)
__global__ void __launch_bounds__(256) k_quotient_combine(const uint64_t* __restrict__ parts, const uint32_t* __restrict__ part_list,
                                                           uint32_t n_direct, const uint64_t* __restrict__ ext,
                                                           const uint32_t* __restrict__ ext_list, uint32_t n_ext,
                                                           const Gate* __restrict__ gates, const uint64_t* __restrict__ cs_lde,
                                                           size_t N, uint32_t num_selectors, uint32_t n_chal, uint32_t log_lde,
                                                           uint32_t mdb, const uint64_t* __restrict__ zh, uint64_t* __restrict__ out) {
  const size_t lde_size = (size_t)1 << log_lde;
  const size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= lde_size) return;
  const size_t i = ntt2::brev((uint32_t)leaf, log_lde);
  uint64_t acc[MAX_CHALLENGES];
#pragma unroll
  for (int c = 0; c < MAX_CHALLENGES; c++) acc[c] = 0;
  for (uint32_t p = 0; p < n_direct; p++) {
    const uint64_t* src = parts + (size_t)part_list[p] * n_chal * lde_size + leaf;
#pragma unroll
    for (int c = 0; c < MAX_CHALLENGES; c++)
      if (c < (int)n_chal) acc[c] = fadd(acc[c], src[(size_t)c * lde_size]);
  }
  for (uint32_t k = 0; k < n_ext; k++) {
    const Gate gate = gates[ext_list[k]];
    const uint64_t filter = gate_filter(gate, gl::canon(cs_lde[(size_t)gate.selector_index * N + leaf]), num_selectors);
    const uint64_t* src = ext + (size_t)k * n_chal * lde_size + leaf;
#pragma unroll
    for (int c = 0; c < MAX_CHALLENGES; c++)
      if (c < (int)n_chal) acc[c] = fadd(acc[c], fmul(filter, src[(size_t)c * lde_size]));
  }
  const uint64_t z_h_inv = zh[((size_t)1 << mdb) + (i & (((size_t)1 << mdb) - 1))];
#pragma unroll
  for (int c = 0; c < MAX_CHALLENGES; c++)
    if (c < (int)n_chal) out[(size_t)c * lde_size + i] = fmul(acc[c], z_h_inv);
}

// apow[c * n_terms + k] = alphas[c]^k
__global__ void k_build_apow(const uint64_t* __restrict__ alphas, uint32_t n_terms, uint64_t* __restrict__ apow) {
  const uint32_t c = blockIdx.x;
  if (threadIdx.x != 0) return;
  const uint64_t a = gl::canon(alphas[c]);
  uint64_t p = 1;
  for (uint32_t k = 0; k < n_terms; k++) {
    apow[(size_t)c * n_terms + k] = p;
    p = fmul(p, a);
  }
}

// apow_gates[q][c] = apow[c][first_gate_term + q] (zero for c >= n_chal)
__global__ void __launch_bounds__(256) k_interleave_apow(const uint64_t* __restrict__ apow, uint32_t n_terms, uint32_t first_gate_term,
                                                          uint32_t n_chal, uint64_t* __restrict__ out) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (first_gate_term + q >= n_terms) return;
#pragma unroll
  for (int c = 0; c < MAX_CHALLENGES; c++) out[(size_t)q * MAX_CHALLENGES + c] = c < (int)n_chal ? apow[(size_t)c * n_terms + first_gate_term + q] : 0;
}

// coefficient k of every column *= base^k  (the second half of coset_ifft: divide by shift^k)
__global__ void __launch_bounds__(256) k_scale_by_powers(uint64_t* __restrict__ data, size_t len, uint64_t base) {
  const size_t k0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (k0 >= len) return;
  uint64_t* col = data + (size_t)blockIdx.y * len;
  uint64_t p = gl::pow(base, k0);
  for (size_t k = k0; k < k0 + 16 && k < len; k++) {
    col[k] = fmul(col[k], p);
    p = fmul(p, base);
  }
}

}  // namespace plonk
