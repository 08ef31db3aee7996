/* plonk.c — CPU restatement of the PLONK-side prover stages between the commitments.  TEST INFRASTRUCTURE.
 *
 * Follows the published plonky2 0.2.2 algorithms (un-vendored dependency plonky2-hwa, Cargo.toml:101-102;
 * SURVEY.md §3.2, A.8, §8 rows a6-a8):
 *   plonk/prover.rs      wires_permutation_partial_products_and_zs, compute_quotient_polys
 *   plonk/vanishing_poly.rs  eval_vanishing_poly_base_batch, evaluate_gate_constraints_base_batch
 *   plonk/plonk_common.rs    ZeroPolyOnCoset, reduce_with_powers_multi, check_partial_products
 *   gates/gate.rs        eval_filtered_base_batch / compute_filter (selector groups, UNUSED_SELECTOR)
 *   gates/{noop,constant,public_input,arithmetic_base,poseidon,base_sum}.rs   eval_unfiltered_base_one
 *   gates/{arithmetic_extension,multiplication_extension,reducing,reducing_extension,random_access,
 *          poseidon_mds,coset_interpolation}.rs                                eval_unfiltered_base_one
 * and, for the gates that live in the reference tree itself, the reference's scalar eval_unfiltered:
 *   city_common_circuit/src/u32/gates/arithmetic_u32.rs:88-150  (U32ArithmeticGate)
 *   city_common_circuit/src/u32/gates/add_many_u32.rs:87-135    (U32AddManyGate)
 *   city_common_circuit/src/u32/gates/subtraction_u32.rs:82-125 (U32SubtractionGate)
 *   city_common_circuit/src/u32/gates/range_check_u32.rs:51-75  (U32RangeCheckGate)
 *   city_common_circuit/src/u32/gates/interleave_u32.rs:86-127  (U32InterleaveGate)
 *   city_common_circuit/src/u32/gates/uninterleave_to_u32.rs:93-136, uninterleave_to_b32.rs:97-141
 *   city_common_circuit/src/u32/gates/comparison.rs:96-170      (ComparisonGate; (32, 16) at builder/pad_circuit.rs:33)
 * Parameters a circuit is described by follow city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145.
 * PARITY STATUS: unpinned by reference fixtures; checked in tests/ against an independent
 * extension-field evaluation of the verifier identity vanishing(zeta) = Z_H(zeta) * t(zeta).
 *
 * Deliberately written the way plonky2 does it (materialise every constraint, then reduce_with_powers), not
 * the way the CUDA kernel does (streaming accumulation), so that the two are independent. */
#include <stdlib.h>
#include <string.h>

#include "gl_inline.h"
#include "p2oracle.h"

#include "poseidon_fast.inc"
static const uint64_t PRC[360] = {
#include "poseidon_rc.inc"
};
static const uint64_t MDS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#define UNUSED_SELECTOR 0xFFFFFFFFull

static inline uint64_t fneg(uint64_t a) { return a ? GL_P - a : 0; }
static uint64_t finv(uint64_t a) { return gl_inv(a); }
static void batch_inverse(const uint64_t *x, uint64_t *out, size_t n) {
  /* Montgomery's trick (plonky2_field batch_multiplicative_inverse computes the same values) */
  uint64_t *pre = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t acc = 1;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = gli_mul(acc, x[i]);
  }
  uint64_t inv = finv(acc);
  for (size_t i = n; i-- > 0;) {
    out[i] = gli_mul(inv, pre[i]);
    inv = gli_mul(inv, x[i]);
  }
  free(pre);
}

/* ---------------------------------------------------------------- partial products and Z (a6) */
void plonk_partial_products_and_zs(const p2o_circuit *c, const uint64_t *wires, const uint64_t *sigmas,
                                   const uint64_t *betas, const uint64_t *gammas, uint64_t *out) {
  const size_t n = (size_t)1 << c->degree_bits;
  const unsigned nr = c->num_routed_wires, deg = c->quotient_degree_factor, npp = c->num_partial_products;
  const unsigned nch = c->num_challenges, nchunks = npp + 1;
  const uint64_t w = gl_root_of_unity(c->degree_bits);
  uint64_t *subgroup = (uint64_t *)malloc(n * sizeof(uint64_t));
  subgroup[0] = 1;
  for (size_t i = 1; i < n; i++) subgroup[i] = gli_mul(subgroup[i - 1], w);
  for (unsigned ch = 0; ch < nch; ch++) {
    const uint64_t beta = gl_canon(betas[ch]), gamma = gl_canon(gammas[ch]);
    uint64_t *chunk_products = (uint64_t *)malloc(n * nchunks * sizeof(uint64_t));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
      uint64_t num[256], den[256], den_inv[256];
      const uint64_t x = subgroup[i];
      for (unsigned j = 0; j < nr; j++) {
        uint64_t wv = gl_canon(wires[(size_t)j * n + i]);
        uint64_t s_id = gli_mul(gl_canon(c->k_is[j]), x);
        num[j] = gli_add(gli_add(wv, gli_mul(beta, s_id)), gamma);
        den[j] = gli_add(gli_add(wv, gli_mul(beta, gl_canon(sigmas[(size_t)j * n + i]))), gamma);
      }
      batch_inverse(den, den_inv, nr);
      for (unsigned k = 0; k < nchunks; k++) {
        uint64_t p = 1;
        for (unsigned j = k * deg; j < (k + 1) * deg && j < nr; j++) p = gli_mul(p, gli_mul(num[j], den_inv[j]));
        chunk_products[i * nchunks + k] = p;
      }
    }
    /* running product: row i holds [pp_0 .. pp_{npp-1}, Z(x_i)] with Z(x_0) = 1 */
    uint64_t z = 1;
    uint64_t *zcol = out + (size_t)ch * n;
    uint64_t *pp = out + (size_t)(nch + ch * npp) * n;
    for (size_t i = 0; i < n; i++) {
      zcol[i] = z;
      uint64_t acc = z;
      for (unsigned k = 0; k < nchunks; k++) {
        acc = gli_mul(acc, chunk_products[i * nchunks + k]);
        if (k < npp) pp[(size_t)k * n + i] = acc;
      }
      z = acc;
    }
    free(chunk_products);
  }
  free(subgroup);
}

/* ---------------------------------------------------------------- gates (a8) */
typedef struct {
  const uint64_t *consts; /* local_constants after the selector prefix */
  const uint64_t *wires;
  const uint64_t *pi_hash;
} gate_vars;

static inline uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gli_mul(x, x), x4 = gli_mul(x2, x2), x3 = gli_mul(x, x2);
  return gli_mul(x3, x4);
}
static void mds_layer(uint64_t s[12]) {
  uint64_t o[12];
  for (int r = 0; r < 12; r++) {
    uint64_t acc = 0;
    for (int i = 0; i < 12; i++) acc = gli_add(acc, gli_mul(s[(i + r) % 12], MDS_CIRC[i]));
    if (r == 0) acc = gli_add(acc, gli_mul(s[0], 8));
    o[r] = acc;
  }
  memcpy(s, o, sizeof o);
}

/* extension elements held in two wires: (a0 + a1 X)(b0 + b1 X), X^2 = 7 */
static inline void e2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t o[2]) {
  uint64_t c0 = gli_add(gli_mul(a[0], b[0]), gli_mul(7, gli_mul(a[1], b[1])));
  uint64_t c1 = gli_add(gli_mul(a[0], b[1]), gli_mul(a[1], b[0]));
  o[0] = c0;
  o[1] = c1;
}

/* returns the number of constraints written to `out` */
static unsigned eval_gate(const p2o_gate *g, const gate_vars *v, uint64_t *out) {
  unsigned k = 0;
  const uint64_t *w = v->wires;
  switch (g->kind) {
    case P2O_GATE_NOOP:
      return 0;
    case P2O_GATE_CONSTANT:
      for (unsigned i = 0; i < g->p0; i++) out[k++] = gli_sub(v->consts[i], w[i]);
      return k;
    case P2O_GATE_PUBLIC_INPUT:
      for (unsigned i = 0; i < 4; i++) out[k++] = gli_sub(w[i], v->pi_hash[i]);
      return k;
    case P2O_GATE_ARITHMETIC:
      for (unsigned i = 0; i < g->p0; i++) {
        uint64_t m0 = w[4 * i], m1 = w[4 * i + 1], ad = w[4 * i + 2], o = w[4 * i + 3];
        uint64_t computed = gli_add(gli_mul(gli_mul(m0, m1), v->consts[0]), gli_mul(ad, v->consts[1]));
        out[k++] = gli_sub(o, computed);
      }
      return k;
    case P2O_GATE_BASE_SUM: { /* BaseSumGate<2>: wire 0 = sum, wires 1.. = limbs (little endian) */
      uint64_t sum = 0;
      for (unsigned i = g->p0; i-- > 0;) sum = gli_add(gli_mul(sum, 2), w[1 + i]);
      out[k++] = gli_sub(sum, w[0]);
      for (unsigned i = 0; i < g->p0; i++) out[k++] = gli_mul(w[1 + i], gli_sub(w[1 + i], 1));
      return k;
    }
    case P2O_GATE_POSEIDON: {
      enum { SWAP = 24, DELTA = 25, FULL0 = 29, PARTIAL = 65, FULL1 = 87 };
      uint64_t swap = w[SWAP];
      out[k++] = gli_mul(swap, gli_sub(swap, 1));
      for (int i = 0; i < 4; i++) out[k++] = gli_sub(gli_mul(swap, gli_sub(w[i + 4], w[i])), w[DELTA + i]);
      uint64_t s[12];
      for (int i = 0; i < 4; i++) {
        s[i] = gli_add(w[i], w[DELTA + i]);
        s[i + 4] = gli_sub(w[i + 4], w[DELTA + i]);
      }
      for (int i = 8; i < 12; i++) s[i] = w[i];
      unsigned round = 0;
      for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], PRC[12 * round + i]);
        if (r != 0)
          for (int i = 0; i < 12; i++) {
            uint64_t in = w[FULL0 + 12 * (r - 1) + i];
            out[k++] = gli_sub(s[i], in);
            s[i] = in;
          }
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
        round++;
      }
      for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], PFAST_FIRST[i]);
      { /* mds_partial_layer_init */
        uint64_t o[12];
        o[0] = s[0];
        for (int i = 0; i < 11; i++) {
          uint64_t acc = 0;
          for (int j = 0; j < 11; j++) acc = gli_add(acc, gli_mul(PFAST_INIT[i * 11 + j], s[1 + j]));
          o[1 + i] = acc;
        }
        memcpy(s, o, sizeof o);
      }
      for (int r = 0; r < 22; r++) {
        uint64_t in = w[PARTIAL + r];
        out[k++] = gli_sub(s[0], in);
        s[0] = sbox7(in);
        if (r < 21) s[0] = gli_add(s[0], PFAST_POST[r]);
        uint64_t d = gli_mul(s[0], 25);
        for (int i = 1; i < 12; i++) d = gli_add(d, gli_mul(s[i], PFAST_W_HATS[r * 11 + i - 1]));
        for (int i = 1; i < 12; i++) s[i] = gli_add(s[i], gli_mul(s[0], PFAST_VS[r * 11 + i - 1]));
        s[0] = d;
      }
      round += 22;
      for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = gli_add(s[i], PRC[12 * round + i]);
        for (int i = 0; i < 12; i++) {
          uint64_t in = w[FULL1 + 12 * r + i];
          out[k++] = gli_sub(s[i], in);
          s[i] = in;
        }
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
        mds_layer(s);
        round++;
      }
      for (int i = 0; i < 12; i++) out[k++] = gli_sub(s[i], w[12 + i]);
      return k;
    }
    case P2O_GATE_U32_ARITHMETIC: { /* arithmetic_u32.rs:88-150; 6 routed wires per op, 32 2-bit limbs */
      const unsigned ops = g->p0;
      for (unsigned i = 0; i < ops; i++) {
        uint64_t m0 = w[6 * i], m1 = w[6 * i + 1], ad = w[6 * i + 2];
        uint64_t lo = w[6 * i + 3], hi = w[6 * i + 4], inv = w[6 * i + 5];
        uint64_t computed = gli_add(gli_mul(m0, m1), ad);
        uint64_t diff = gli_sub(0xFFFFFFFFull, hi);
        uint64_t hi_not_max = gli_sub(gli_mul(inv, diff), 1);
        out[k++] = gli_mul(hi_not_max, lo);
        out[k++] = gli_sub(gli_add(gli_mul(hi, 1ull << 32), lo), computed);
        uint64_t clo = 0, chi = 0;
        for (int j = 31; j >= 0; j--) {
          uint64_t limb = w[6 * ops + 32 * i + j];
          uint64_t prod = 1;
          for (uint64_t x = 0; x < 4; x++) prod = gli_mul(prod, gli_sub(limb, x));
          out[k++] = prod;
          if (j < 16)
            clo = gli_add(gli_mul(clo, 4), limb);
          else
            chi = gli_add(gli_mul(chi, 4), limb);
        }
        out[k++] = gli_sub(clo, lo);
        out[k++] = gli_sub(chi, hi);
      }
      return k;
    }
    case P2O_GATE_U32_ADD_MANY: { /* add_many_u32.rs:87-135; p0 = num_addends, p1 = num_ops; 16 + 2 limbs */
      const unsigned na = g->p0, ops = g->p1, per = na + 3;
      for (unsigned i = 0; i < ops; i++) {
        uint64_t computed = 0;
        for (unsigned j = 0; j < na; j++) computed = gli_add(computed, w[per * i + j]);
        computed = gli_add(computed, w[per * i + na]);
        uint64_t res = w[per * i + na + 1], carry = w[per * i + na + 2];
        out[k++] = gli_sub(gli_add(gli_mul(carry, 1ull << 32), res), computed);
        uint64_t cres = 0, ccar = 0;
        for (int j = 17; j >= 0; j--) {
          uint64_t limb = w[per * ops + 18 * i + j];
          uint64_t prod = 1;
          for (uint64_t x = 0; x < 4; x++) prod = gli_mul(prod, gli_sub(limb, x));
          out[k++] = prod;
          if (j < 16)
            cres = gli_add(gli_mul(cres, 4), limb);
          else
            ccar = gli_add(gli_mul(ccar, 4), limb);
        }
        out[k++] = gli_sub(cres, res);
        out[k++] = gli_sub(ccar, carry);
      }
      return k;
    }
    case P2O_GATE_U32_SUBTRACTION: { /* subtraction_u32.rs:82-125; 5 routed wires per op, 16 2-bit limbs */
      const unsigned ops = g->p0;
      for (unsigned i = 0; i < ops; i++) {
        uint64_t x = w[5 * i], y = w[5 * i + 1], bin = w[5 * i + 2], res = w[5 * i + 3], bout = w[5 * i + 4];
        uint64_t initial = gli_sub(gli_sub(x, y), bin);
        out[k++] = gli_sub(res, gli_add(initial, gli_mul(bout, 1ull << 32)));
        uint64_t comb = 0;
        for (int j = 15; j >= 0; j--) {
          uint64_t limb = w[5 * ops + 16 * i + j];
          uint64_t prod = 1;
          for (uint64_t t = 0; t < 4; t++) prod = gli_mul(prod, gli_sub(limb, t));
          out[k++] = prod;
          comb = gli_add(gli_mul(comb, 4), limb);
        }
        out[k++] = gli_sub(comb, res);
        out[k++] = gli_mul(bout, gli_sub(1, bout));
      }
      return k;
    }
    case P2O_GATE_U32_RANGE_CHECK: { /* range_check_u32.rs:51-75; p0 = num_input_limbs, 16 aux 2-bit limbs each */
      const unsigned nl = g->p0;
      for (unsigned i = 0; i < nl; i++) {
        uint64_t comb = 0;
        for (int j = 15; j >= 0; j--) comb = gli_add(gli_mul(comb, 4), w[nl + 16 * i + j]);
        out[k++] = gli_sub(comb, w[i]);
        for (int j = 0; j < 16; j++) {
          uint64_t limb = w[nl + 16 * i + j];
          uint64_t prod = 1;
          for (uint64_t t = 0; t < 4; t++) prod = gli_mul(prod, gli_sub(limb, t));
          out[k++] = prod;
        }
      }
      return k;
    }
    case P2O_GATE_U32_INTERLEAVE: { /* p0 = num_ops; 2 routed wires per op, 32 big-endian bits */
      const unsigned ops = g->p0;
      for (unsigned i = 0; i < ops; i++) {
        const uint64_t *bits = w + 2 * ops + 32 * i;
        uint64_t cx = 0, cxi = 0;
        for (int j = 0; j < 32; j++) { /* reduce_with_powers(bits.rev(), base) */
          cx = gli_add(gli_mul(cx, 2), bits[j]);
          cxi = gli_add(gli_mul(cxi, 4), bits[j]);
        }
        out[k++] = gli_sub(cx, w[2 * i]);
        out[k++] = gli_sub(cxi, w[2 * i + 1]);
        for (int j = 0; j < 32; j++) out[k++] = gli_mul(bits[j], gli_sub(bits[j], 1));
      }
      return k;
    }
    case P2O_GATE_UNINTERLEAVE_TO_U32:
    case P2O_GATE_UNINTERLEAVE_TO_B32: { /* p0 = num_ops; 3 routed wires per op, 64 big-endian bits */
      const unsigned ops = g->p0;
      const int b32 = g->kind == P2O_GATE_UNINTERLEAVE_TO_B32;
      for (unsigned i = 0; i < ops; i++) {
        const uint64_t *bits = w + 3 * ops + 64 * i;
        uint64_t cx = 0, ce = 0, co = 0;
        for (int j = 0; j < 64; j++) cx = gli_add(gli_mul(cx, 2), bits[j]);
        out[k++] = gli_sub(cx, w[3 * i]);
        for (int j = 0; j < 32; j++) {
          uint64_t coeff = b32 ? (1ull << (2 * (31 - j))) : (1ull << (31 - j));
          ce = gli_add(ce, gli_mul(coeff, bits[2 * j]));
          co = gli_add(co, gli_mul(coeff, bits[2 * j + 1]));
        }
        out[k++] = gli_sub(ce, w[3 * i + 1]);
        out[k++] = gli_sub(co, w[3 * i + 2]);
        for (int j = 0; j < 64; j++) out[k++] = gli_mul(bits[j], gli_sub(bits[j], 1));
      }
      return k;
    }
    case P2O_GATE_COMPARISON: { /* p0 = num_bits, p1 = num_chunks */
      const unsigned nc = g->p1, cb = (g->p0 + nc - 1) / nc;
      const uint64_t *fc = w + 4, *sc = w + 4 + nc;
      uint64_t f_comb = 0, s_comb = 0;
      for (unsigned i = nc; i-- > 0;) {
        f_comb = gli_add(gli_mul(f_comb, 1ull << cb), fc[i]);
        s_comb = gli_add(gli_mul(s_comb, 1ull << cb), sc[i]);
      }
      out[k++] = gli_sub(f_comb, w[0]);
      out[k++] = gli_sub(s_comb, w[1]);
      uint64_t msd = 0;
      for (unsigned i = 0; i < nc; i++) {
        uint64_t fp = 1, sp = 1;
        for (uint64_t x = 0; x < (1ull << cb); x++) {
          fp = gli_mul(fp, gli_sub(fc[i], x));
          sp = gli_mul(sp, gli_sub(sc[i], x));
        }
        out[k++] = fp;
        out[k++] = sp;
        uint64_t diff = gli_sub(sc[i], fc[i]);
        uint64_t dummy = w[4 + 2 * nc + i], eq = w[4 + 3 * nc + i], inter = w[4 + 4 * nc + i];
        out[k++] = gli_sub(gli_mul(diff, dummy), gli_sub(1, eq));
        out[k++] = gli_mul(eq, diff);
        out[k++] = gli_sub(inter, gli_mul(eq, msd));
        msd = gli_add(inter, gli_mul(gli_sub(1, eq), diff));
      }
      out[k++] = gli_sub(w[3], msd);
      const uint64_t *bits = w + 4 + 5 * nc;
      uint64_t comb = 0;
      for (unsigned b = 0; b <= cb; b++) out[k++] = gli_mul(bits[b], gli_sub(1, bits[b]));
      for (unsigned b = cb + 1; b-- > 0;) comb = gli_add(gli_mul(comb, 2), bits[b]);
      out[k++] = gli_sub(gli_add(1ull << cb, w[3]), comb);
      out[k++] = gli_sub(w[2], bits[cb]);
      return k;
    }
    case P2O_GATE_ARITHMETIC_EXT:
    case P2O_GATE_MUL_EXT: { /* p0 = num_ops; output - (m0 * m1 * c0 [+ addend * c1]) componentwise */
      const int arith = g->kind == P2O_GATE_ARITHMETIC_EXT;
      const unsigned per = arith ? 8 : 6;
      for (unsigned i = 0; i < g->p0; i++) {
        const uint64_t *q = w + per * i;
        uint64_t prod[2];
        e2_mul(q, q + 2, prod);
        for (int t = 0; t < 2; t++) {
          uint64_t computed = gli_mul(prod[t], v->consts[0]);
          if (arith) computed = gli_add(computed, gli_mul(q[4 + t], v->consts[1]));
          out[k++] = gli_sub(q[per - 2 + t], computed);
        }
      }
      return k;
    }
    case P2O_GATE_REDUCING:
    case P2O_GATE_REDUCING_EXT: { /* p0 = num_coeffs; output 0..2, alpha 2..4, old_acc 4..6, coeffs from 6 */
      const unsigned n = g->p0;
      const int ext = g->kind == P2O_GATE_REDUCING_EXT;
      const unsigned start_accs = 6 + (ext ? 2 * n : n);
      uint64_t acc[2] = {w[4], w[5]};
      for (unsigned i = 0; i < n; i++) {
        uint64_t t[2];
        e2_mul(acc, w + 2, t);
        const uint64_t *nxt = i == n - 1 ? w : w + start_accs + 2 * i;
        uint64_t c0 = ext ? w[6 + 2 * i] : w[6 + i], c1 = ext ? w[7 + 2 * i] : 0;
        out[k++] = gli_sub(gli_add(t[0], c0), nxt[0]);
        out[k++] = gli_sub(gli_add(t[1], c1), nxt[1]);
        acc[0] = nxt[0];
        acc[1] = nxt[1];
      }
      return k;
    }
    case P2O_GATE_RANDOM_ACCESS: { /* p0 = bits, p1 = num_copies | num_extra_constants << 16 */
      const unsigned bits = g->p0, copies = g->p1 & 0xFFFF, extra = g->p1 >> 16, vec = 1u << bits;
      const unsigned routed = (2 + vec) * copies + extra;
      for (unsigned cp = 0; cp < copies; cp++) {
        const uint64_t *q = w + (2 + vec) * cp, *b = w + routed + cp * bits;
        uint64_t rec = 0, items[64];
        for (unsigned t = 0; t < bits; t++) out[k++] = gli_mul(b[t], gli_sub(b[t], 1));
        for (unsigned t = bits; t-- > 0;) rec = gli_add(gli_add(rec, rec), b[t]);
        out[k++] = gli_sub(rec, q[0]);
        for (unsigned t = 0; t < vec; t++) items[t] = q[2 + t];
        unsigned len = vec;
        for (unsigned t = 0; t < bits; t++) {
          len >>= 1;
          for (unsigned u = 0; u < len; u++)
            items[u] = gli_add(items[2 * u], gli_mul(b[t], gli_sub(items[2 * u + 1], items[2 * u])));
        }
        out[k++] = gli_sub(items[0], q[1]);
      }
      for (unsigned t = 0; t < extra; t++) out[k++] = gli_sub(v->consts[t], w[(2 + vec) * copies + t]);
      return k;
    }
    case P2O_GATE_POSEIDON_MDS: { /* inputs 12 x 2 wires, outputs 12 x 2 wires: the MDS layer componentwise */
      for (int r = 0; r < 12; r++)
        for (int t = 0; t < 2; t++) {
          uint64_t acc = 0;
          for (int i = 0; i < 12; i++) acc = gli_add(acc, gli_mul(w[2 * ((i + r) % 12) + t], MDS_CIRC[i]));
          if (r == 0) acc = gli_add(acc, gli_mul(w[t], 8));
          out[k++] = gli_sub(acc, w[24 + 2 * r + t]);
        }
      return k;
    }
    case P2O_GATE_COSET_INTERPOLATION: { /* p0 = subgroup_bits, p1 = degree (with_max_degree(4, 8) -> 6) */
      const unsigned n = 1u << g->p0, degree = g->p1, n_int = (n - 2) / (degree - 1);
      const unsigned pt = 1 + 2 * n, val = pt + 2, ie0 = pt + 4, ip0 = ie0 + 2 * n_int, sh = ie0 + 4 * n_int;
      uint64_t dom[64], wts[64];
      /* barycentric weights of the subgroup: 1 / prod_{j != i} (x_i - x_j) = x_i / n */
      uint64_t gen = gl_root_of_unity(g->p0), ninv = gl_inv(n);
      dom[0] = 1;
      for (unsigned i = 1; i < n; i++) dom[i] = gli_mul(dom[i - 1], gen);
      for (unsigned i = 0; i < n; i++) wts[i] = gli_mul(dom[i], ninv);
      const uint64_t shift = w[0];
      out[k++] = gli_sub(w[pt], gli_mul(w[sh], shift));
      out[k++] = gli_sub(w[pt + 1], gli_mul(w[sh + 1], shift));
      uint64_t ev[2] = {0, 0}, pr[2] = {1, 0};
      unsigned lo = 0, hi = degree;
      for (unsigned c = 0; c <= n_int; c++) {
        for (unsigned i = lo; i < hi && i < n; i++) { /* partial_interpolate */
          uint64_t term[2] = {gli_sub(w[sh], dom[i]), w[sh + 1]};
          uint64_t wv[2] = {gli_mul(w[1 + 2 * i], wts[i]), gli_mul(w[2 + 2 * i], wts[i])};
          uint64_t a[2], b[2];
          e2_mul(ev, term, a);
          e2_mul(wv, pr, b);
          ev[0] = gli_add(a[0], b[0]);
          ev[1] = gli_add(a[1], b[1]);
          e2_mul(pr, term, a);
          pr[0] = a[0];
          pr[1] = a[1];
        }
        if (c == n_int) break;
        out[k++] = gli_sub(w[ie0 + 2 * c], ev[0]);
        out[k++] = gli_sub(w[ie0 + 2 * c + 1], ev[1]);
        out[k++] = gli_sub(w[ip0 + 2 * c], pr[0]);
        out[k++] = gli_sub(w[ip0 + 2 * c + 1], pr[1]);
        ev[0] = w[ie0 + 2 * c], ev[1] = w[ie0 + 2 * c + 1];
        pr[0] = w[ip0 + 2 * c], pr[1] = w[ip0 + 2 * c + 1];
        lo = 1 + (degree - 1) * (c + 1);
        hi = lo + degree - 1;
      }
      out[k++] = gli_sub(w[val], ev[0]);
      out[k++] = gli_sub(w[val + 1], ev[1]);
      return k;
    }
    default:
      return 0;
  }
}

/* unfiltered constraints of one gate at one point (test hook: compared gate by gate with the Python restatement) */
unsigned plonk_eval_gate(const p2o_gate *g, const uint64_t *wires, const uint64_t *consts, const uint64_t *pi_hash,
                         uint64_t *out) {
  gate_vars v = {consts, wires, pi_hash};
  return eval_gate(g, &v, out);
}

/* gates/gate.rs compute_filter */
static uint64_t compute_filter(const p2o_gate *g, uint64_t s, int many_selectors) {
  uint64_t f = 1;
  for (unsigned i = g->group_start; i < g->group_end; i++)
    if (i != g->row) f = gli_mul(f, gli_sub(i, s));
  if (many_selectors) f = gli_mul(f, gli_sub(UNUSED_SELECTOR, s));
  return f;
}

/* One point: vanishing-polynomial terms reduced with the powers of every alpha (eval_vanishing_poly_base_batch
 * for a batch of one).  x is the coset point, z_h = Z_H(x). */
static void eval_vanishing_point(const p2o_circuit *c, uint64_t x, uint64_t z_h, const uint64_t *consts_sigmas,
                                 const uint64_t *wires, const uint64_t *zs_local, const uint64_t *zs_next,
                                 const uint64_t *pi_hash, const uint64_t *betas, const uint64_t *gammas,
                                 const uint64_t *alphas, uint64_t *res) {
  const unsigned nr = c->num_routed_wires, nch = c->num_challenges, npp = c->num_partial_products;
  const unsigned deg = c->quotient_degree_factor, ngc = c->num_gate_constraints;
  const unsigned n_terms = nch + nch * (npp + 1) + ngc;
  uint64_t *terms = (uint64_t *)calloc(n_terms, sizeof(uint64_t));
  const uint64_t *local_constants = consts_sigmas, *s_sigmas = consts_sigmas + c->num_constants;
  /* L_0(x) = Z_H(x) / (n (x - 1)) */
  uint64_t n_f = (uint64_t)1 << c->degree_bits;
  uint64_t l0 = gli_mul(z_h, finv(gli_mul(n_f, gli_sub(x, 1))));
  for (unsigned i = 0; i < nch; i++) terms[i] = gli_mul(l0, gli_sub(zs_local[i], 1));
  unsigned t = nch;
  for (unsigned i = 0; i < nch; i++) {
    uint64_t num[256], den[256];
    for (unsigned j = 0; j < nr; j++) {
      uint64_t s_id = gli_mul(gl_canon(c->k_is[j]), x);
      num[j] = gli_add(gli_add(wires[j], gli_mul(betas[i], s_id)), gammas[i]);
      den[j] = gli_add(gli_add(wires[j], gli_mul(betas[i], s_sigmas[j])), gammas[i]);
    }
    /* check_partial_products: accumulators z_x, pp_0.., z_gx */
    for (unsigned k = 0; k <= npp; k++) {
      uint64_t prev = k == 0 ? zs_local[i] : zs_local[nch + i * npp + k - 1];
      uint64_t next = k == npp ? zs_next[i] : zs_local[nch + i * npp + k];
      uint64_t np = 1, dp = 1;
      for (unsigned j = k * deg; j < (k + 1) * deg && j < nr; j++) {
        np = gli_mul(np, num[j]);
        dp = gli_mul(dp, den[j]);
      }
      terms[t++] = gli_sub(gli_mul(prev, np), gli_mul(next, dp));
    }
  }
  /* evaluate_gate_constraints_base_batch: every gate adds its filtered constraints into slots 0.. */
  uint64_t *gate_terms = terms + t;
  uint64_t *tmp = (uint64_t *)malloc((ngc + 1) * sizeof(uint64_t));
  gate_vars gv = {local_constants + c->num_selectors, wires, pi_hash};
  for (unsigned g = 0; g < c->n_gates; g++) {
    const p2o_gate *gate = &c->gates[g];
    uint64_t filter = compute_filter(gate, local_constants[gate->selector_index], c->num_selectors > 1);
    unsigned k = eval_gate(gate, &gv, tmp);
    for (unsigned q = 0; q < k && q < ngc; q++) gate_terms[q] = gli_add(gate_terms[q], gli_mul(tmp[q], filter));
  }
  free(tmp);
  /* reduce_with_powers_multi: sum_k term_k alpha^k, Horner from the last term */
  for (unsigned i = 0; i < nch; i++) {
    uint64_t acc = 0;
    for (unsigned q = n_terms; q-- > 0;) acc = gli_add(gli_mul(acc, alphas[i]), terms[q]);
    res[i] = acc;
  }
  free(terms);
}

/* ---------------------------------------------------------------- compute_quotient_polys (a7) */
void plonk_compute_quotient_polys(const p2o_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                                  const uint64_t *wires_leaves, const uint64_t *zs_leaves, const uint64_t *pi_hash_in,
                                  const uint64_t *betas_in, const uint64_t *gammas_in, const uint64_t *alphas_in,
                                  uint64_t *out_chunks) {
  const unsigned nch = c->num_challenges, npp = c->num_partial_products;
  unsigned max_degree_bits = 0;
  while ((1u << max_degree_bits) < c->quotient_degree_factor) max_degree_bits++;
  const unsigned step_bits = rate_bits - max_degree_bits;
  const unsigned log_lde = c->degree_bits + max_degree_bits, log_full = c->degree_bits + rate_bits;
  const size_t lde_size = (size_t)1 << log_lde, n = (size_t)1 << c->degree_bits;
  const size_t next_step = (size_t)1 << max_degree_bits;
  const size_t w_cs = c->num_constants + c->num_routed_wires, w_wires = c->num_wires, w_zs = nch * (1 + npp);
  uint64_t betas[8], gammas[8], alphas[8], pi_hash[4];
  for (unsigned i = 0; i < nch; i++) {
    betas[i] = gl_canon(betas_in[i]);
    gammas[i] = gl_canon(gammas_in[i]);
    alphas[i] = gl_canon(alphas_in[i]);
  }
  for (int i = 0; i < 4; i++) pi_hash[i] = gl_canon(pi_hash_in[i]);
  /* ZeroPolyOnCoset: Z_H(7 w^i) = 7^n * w_{2^max_degree_bits}^(i mod 2^max_degree_bits) - 1 */
  uint64_t g_pow_n = gl_pow(7, n);
  uint64_t zh[256], zh_inv[256];
  uint64_t wr = gl_root_of_unity(max_degree_bits), xr = 1;
  for (size_t i = 0; i < next_step; i++) {
    zh[i] = gli_sub(gli_mul(g_pow_n, xr), 1);
    xr = gli_mul(xr, wr);
  }
  batch_inverse(zh, zh_inv, next_step);
  uint64_t *points = (uint64_t *)malloc(lde_size * sizeof(uint64_t));
  uint64_t wl = gl_root_of_unity(log_lde);
  points[0] = 1;
  for (size_t i = 1; i < lde_size; i++) points[i] = gli_mul(points[i - 1], wl);
  uint64_t *qvals = (uint64_t *)malloc(nch * lde_size * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t i = 0; i < lde_size; i++) {
    const uint64_t x = gli_mul(7, points[i]);
    const size_t i_next = (i + next_step) % lde_size;
    /* get_lde_values(i, step): leaf reverse_bits(i * step, degree_log + rate_bits) */
    const size_t leaf = bitrev(i << step_bits, log_full), leaf_next = bitrev(i_next << step_bits, log_full);
    uint64_t res[8];
    eval_vanishing_point(c, x, zh[i % next_step], cs_leaves + leaf * w_cs, wires_leaves + leaf * w_wires,
                         zs_leaves + leaf * w_zs, zs_leaves + leaf_next * w_zs, pi_hash, betas, gammas, alphas, res);
    for (unsigned ch = 0; ch < nch; ch++) qvals[ch * lde_size + i] = gli_mul(res[ch], zh_inv[i % next_step]);
  }
  /* values.coset_ifft(7), then split into quotient_degree_factor chunks of n coefficients */
  for (unsigned ch = 0; ch < nch; ch++) {
    gl_coset_ifft(qvals + ch * lde_size, log_lde, 7);
    memcpy(out_chunks + (size_t)ch * lde_size, qvals + ch * lde_size, lde_size * sizeof(uint64_t));
  }
  free(points);
  free(qvals);
}
