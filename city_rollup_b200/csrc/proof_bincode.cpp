// proof_bincode.cpp — the byte format the reference stores and ships proofs in.
//
// The worker serialises every `ProofWithPublicInputs<GoldilocksField, PoseidonGoldilocksConfig, 2>` with
// `bincode::serialize` (city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46,65-72; Redis store
// city_redis_store/src/lib.rs:54-82): bincode 1.3.3 defaults = little endian, fixed-width integers, every Vec
// prefixed by its length as a u64, structs and tuples as the plain concatenation of their fields.  p2b_prove
// emits the same fields in the same order as bare u64 words (include/p2b.h); the functions here add / strip the
// length prefixes, so that the bytes a patched `CircuitData::prove` hands to the proof store are produced (and
// child proofs read back) without a Rust-side re-walk of the structure.  Host-only code: formatting, no arithmetic.
//
// Field order (plonky2 0.2.2 plonk/proof.rs, fri/proof.rs; checked on the ten stored proofs of
// qbench_data/example.bin, tests/test_proof_bincode.py):
//   Proof { wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap : MerkleCap = Vec<HashOut>,
//           openings : OpeningSet { constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products,
//                                   quotient_polys, lookup_zs, lookup_zs_next : Vec<Ext> },
//           opening_proof : FriProof { commit_phase_merkle_caps : Vec<MerkleCap>,
//                                      query_round_proofs : Vec<FriQueryRound {
//                                          initial_trees_proof : { evals_proofs : Vec<(Vec<F>, MerkleProof)> },
//                                          steps : Vec<FriQueryStep { evals : Vec<Ext>, merkle_proof }> }>,
//                                      final_poly : PolynomialCoeffs<Ext> = Vec<Ext>, pow_witness : F } },
//   public_inputs : Vec<F>
#include <cstring>
#include <vector>

#include "../../include/p2b.h"

namespace {

struct Walker {
  // one pass over the structure; `emit_len(n)` is called at every Vec boundary, `emit_words(k)` for k payload words
  const p2b_proof_shape& s;
  const p2b_fri_params& fp;
  uint32_t log_N, n_final;
  bool ok;
  Walker(const p2b_proof_shape& s_, const p2b_fri_params& fp_) : s(s_), fp(fp_), log_N(0), n_final(0), ok(false) {
    if (fp.n_layers > P2B_MAX_FRI_LAYERS || s.degree_bits > 40 || fp.rate_bits > 16 || fp.cap_height > 24) return;
    uint32_t sum = 0;
    for (uint32_t l = 0; l < fp.n_layers; l++) {
      if (fp.reduction_arity_bits[l] == 0 || fp.reduction_arity_bits[l] > 16) return;
      sum += fp.reduction_arity_bits[l];
    }
    if (sum > s.degree_bits) return;
    log_N = s.degree_bits + fp.rate_bits;
    if (fp.cap_height > log_N - sum || s.constants_sigmas_cap_height > log_N) return;
    n_final = 1u << (s.degree_bits - sum);
    ok = true;
  }
  template <class L, class W>
  void walk(L&& emit_len, W&& emit_words) const {
    const size_t cap = (size_t)1 << fp.cap_height;
    const size_t nch = s.num_challenges;
    for (int i = 0; i < 3; i++) emit_len(cap), emit_words(4 * cap);
    const size_t open[9] = {s.num_constants, s.num_routed_wires, s.num_wires, nch, nch, nch * s.num_partial_products,
                            nch * s.quotient_degree_factor, 0, 0};
    for (int i = 0; i < 9; i++) emit_len(open[i]), emit_words(2 * open[i]);
    emit_len(fp.n_layers);
    for (uint32_t l = 0; l < fp.n_layers; l++) emit_len(cap), emit_words(4 * cap);
    const size_t width[4] = {(size_t)s.num_constants + s.num_routed_wires, s.num_wires,
                             nch * (1 + s.num_partial_products), nch * s.quotient_degree_factor};
    emit_len(fp.num_query_rounds);
    for (uint32_t q = 0; q < fp.num_query_rounds; q++) {
      emit_len(4);
      for (int o = 0; o < 4; o++) {
        const size_t sib = log_N - (o == 0 ? s.constants_sigmas_cap_height : fp.cap_height);
        emit_len(width[o]), emit_words(width[o]);
        emit_len(sib), emit_words(4 * sib);
      }
      emit_len(fp.n_layers);
      uint32_t log_cur = log_N;
      for (uint32_t l = 0; l < fp.n_layers; l++) {
        const uint32_t ab = fp.reduction_arity_bits[l];
        log_cur -= ab;
        const size_t sib = log_cur - fp.cap_height;
        emit_len((size_t)1 << ab), emit_words((size_t)2 << ab);
        emit_len(sib), emit_words(4 * sib);
      }
    }
    emit_len(n_final), emit_words(2 * (size_t)n_final);
    emit_words(1);  // pow_witness
    emit_len(s.n_public_inputs), emit_words(s.n_public_inputs);
  }
};

inline void put_u64(uint8_t* p, uint64_t v) {
  for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}
inline uint64_t get_u64(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
  return v;
}

}  // namespace

extern "C" size_t p2b_proof_words(const p2b_proof_shape* shape, const p2b_fri_params* params) {
  if (!shape || !params) return 0;
  Walker w(*shape, *params);
  if (!w.ok) return 0;
  size_t words = 0;
  w.walk([](size_t) {}, [&](size_t k) { words += k; });
  return words;
}

extern "C" size_t p2b_proof_bincode_len(const p2b_proof_shape* shape, const p2b_fri_params* params) {
  if (!shape || !params) return 0;
  Walker w(*shape, *params);
  if (!w.ok) return 0;
  size_t words = 0;
  w.walk([&](size_t) { words += 1; }, [&](size_t k) { words += k; });
  return 8 * words;
}

extern "C" int p2b_proof_to_bincode(const p2b_proof_shape* shape, const p2b_fri_params* params, const uint64_t* words,
                                    size_t n_words, uint8_t* out, size_t out_cap, size_t* written) {
  if (!shape || !params || !words || !out) return P2B_ERR_INVALID;
  Walker w(*shape, *params);
  if (!w.ok) return P2B_ERR_INVALID;
  if (n_words != p2b_proof_words(shape, params) || out_cap < p2b_proof_bincode_len(shape, params)) return P2B_ERR_INVALID;
  size_t o = 0, i = 0;
  w.walk([&](size_t n) { put_u64(out + o, n), o += 8; },
         [&](size_t k) {
           for (size_t j = 0; j < k; j++) put_u64(out + o + 8 * j, words[i + j]);
           o += 8 * k, i += k;
         });
  if (written) *written = o;
  return P2B_OK;
}

extern "C" int p2b_proof_from_bincode(const p2b_proof_shape* shape, const p2b_fri_params* params, const uint8_t* bytes,
                                      size_t n_bytes, uint64_t* words_out, size_t words_cap, size_t* n_words) {
  if (!shape || !params || !bytes || !words_out) return P2B_ERR_INVALID;
  Walker w(*shape, *params);
  if (!w.ok) return P2B_ERR_INVALID;
  // the blob must have exactly the shape's length and every length prefix must be the one the shape dictates:
  // a proof of another circuit is rejected here rather than mis-sliced downstream
  if (n_bytes != p2b_proof_bincode_len(shape, params) || words_cap < p2b_proof_words(shape, params)) return P2B_ERR_INVALID;
  size_t o = 0, i = 0;
  bool good = true;
  w.walk([&](size_t n) { good = good && get_u64(bytes + o) == (uint64_t)n, o += 8; },
         [&](size_t k) {
           for (size_t j = 0; j < k; j++) words_out[i + j] = get_u64(bytes + o + 8 * j);
           o += 8 * k, i += k;
         });
  if (!good) return P2B_ERR_INVALID;
  if (n_words) *n_words = i;
  return P2B_OK;
}
