/* p2oracle.h — CPU oracle for the Plonky2 proving hot path behind City Rollup worker jobs.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (city_rollup_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, as the checker / the reported CPU baseline.
 *
 * What it restates: the reference (QEDProtocol/city-rollup) reaches this path only through
 * `circuit_data.prove(pw)` (e.g. city_common_circuit/src/proof_minifier/pm_core.rs:151); the
 * arithmetic lives in the un-vendored git dependency plonky2 0.2.2 @ QEDProtocol/plonky2-hwa
 * rev 6a8ca008 (Cargo.toml:101-102, Cargo.lock:4174-4223), which is NOT on disk.  This file
 * therefore restates the published plonky2 0.2.2 algorithms and is pinned
 * against every known answer the reference tree itself holds for the path:
 *   K1/K2 city_crypto/src/hash/cached_zero_hashes.rs:10-1036,1039-2066 (Poseidon permutation, sponge)
 *   K3    qbench_data/example.bin (10 stored proofs: Merkle leaf/path/cap conventions, FRI folds)
 *   K4/K5 city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145 (generator 7, FRI params)
 * PARITY STATUS: Poseidon / sponge / Merkle / FRI-fold conventions are pinned by those vectors;
 * NTT/LDE values, caps of given polynomials and the transcript are NOT pinned by any reference
 * test or fixture — for those this oracle is "parity unpinned" beyond exactness
 * of field arithmetic and the self-consistency checks in tests/.
 */
#ifndef P2ORACLE_H
#define P2ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */

/* ---- field (plonky2_field GoldilocksField) ---- */
uint64_t gl_canon(uint64_t a);
uint64_t gl_add(uint64_t a, uint64_t b);
uint64_t gl_sub(uint64_t a, uint64_t b);
uint64_t gl_mul(uint64_t a, uint64_t b);
uint64_t gl_pow(uint64_t a, uint64_t e);
uint64_t gl_inv(uint64_t a);
uint64_t gl_root_of_unity(unsigned log_n); /* primitive 2^log_n-th root, plonky2 convention */
/* quadratic extension F[X]/(X^2-7); element = {c0,c1} */
void gl2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
void gl2_inv(const uint64_t a[2], uint64_t out[2]);

/* ---- Poseidon ---- */
void poseidon_permute(uint64_t state[12]);      /* the literal round structure */
void poseidon_permute_fast(uint64_t state[12]); /* fast-partial-round form (what PoseidonGate constrains); same result */
void poseidon_hash_no_pad(const uint64_t *in, size_t len, uint64_t out[4]);
void poseidon_hash_or_noop(const uint64_t *in, size_t len, uint64_t out[4]);
void poseidon_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

/* ---- Merkle tree (plonky2 hash/merkle_tree.rs) ----
 * leaves: row-major n_leaves x leaf_len.  digests_out: 2*(n_leaves - 2^cap_height) digests of 4
 * u64 in plonky2's interleaved layout.  cap_out: 2^cap_height digests. */
void merkle_tree_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len, unsigned cap_height,
                     uint64_t *digests_out, uint64_t *cap_out);
/* siblings_out: (log2(n_leaves)-cap_height) digests, leaf level first. */
void merkle_prove(const uint64_t *digests, size_t n_leaves, unsigned cap_height, size_t leaf_index,
                  uint64_t *siblings_out);
/* returns 1 if the path leads to cap[leaf_index >> n_siblings] */
int merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index, const uint64_t *siblings,
                  unsigned n_siblings, const uint64_t *cap);

/* ---- FFT (plonky2_field fft.rs / polynomial/mod.rs) ---- all natural order ---- */
void gl_fft(uint64_t *a, unsigned log_n);                    /* coeffs -> values, in place */
void gl_ifft(uint64_t *a, unsigned log_n);                   /* values -> coeffs, in place */
void gl_coset_fft(uint64_t *a, unsigned log_n, uint64_t shift);
void gl_coset_ifft(uint64_t *a, unsigned log_n, uint64_t shift);
void gl2_coset_fft(uint64_t *a /* n ext elems, interleaved */, unsigned log_n, uint64_t shift);

/* ---- PolynomialBatch (plonky2 fri/oracle.rs) ----
 * cols: n_cols pointers to 2^log_n u64 each.  Outputs (each may be NULL to skip):
 *   coeffs_out  n_cols x n (column-major, natural order)   [from_values only]
 *   leaves_out  (n<<rate_bits) x n_cols row-major, row j = LDE row bitrev(j) on coset 7*<w>
 *   digests_out 2*((n<<rate_bits) - 2^cap_height) x 4, plonky2 layout;  cap_out 2^cap_height x 4
 * blinding (salted) batches are not used by any worker circuit and not restated. */
void batch_from_coeffs(const uint64_t *const *cols, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       unsigned cap_height, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out);
void batch_from_values(const uint64_t *const *cols, size_t n_cols, unsigned log_n, unsigned rate_bits,
                       unsigned cap_height, uint64_t *coeffs_out, uint64_t *leaves_out,
                       uint64_t *digests_out, uint64_t *cap_out);

/* ---- Challenger (plonky2 iop/challenger.rs) ---- */
typedef struct {
  uint64_t state[12];
  uint64_t in[8];
  uint64_t out[8];
  unsigned n_in, n_out;
} p2o_challenger;
void challenger_init(p2o_challenger *c);
void challenger_observe(p2o_challenger *c, const uint64_t *elems, size_t n);
uint64_t challenger_get(p2o_challenger *c);

/* ---- FRI commit phase + PoW (plonky2 fri/prover.rs) ----
 * coeffs/values: len ext elements (interleaved c0,c1), values in natural order on coset 7*<w_len>.
 * Per layer i (arity 2^arity_bits[i]): the tree is built over bit-reversed values chunked by arity.
 * Outputs: caps_out  n_layers x 2^cap_height x 4;  layer_leaves_out[i] (may be NULL) receives the
 * layer's leaves row-major (len_i/arity x 2*arity); layer_digests_out[i] likewise (plonky2 layout).
 * final_poly_out: (len >> sum(arity_bits) >> rate_bits) ext elements.  The challenger is advanced
 * exactly as fri_committed_trees does (observe cap, squeeze beta, ..., observe final poly). */
void fri_committed_trees(const uint64_t *coeffs, const uint64_t *values, size_t len,
                         const unsigned *arity_bits, size_t n_layers, unsigned rate_bits,
                         unsigned cap_height, p2o_challenger *ch, uint64_t *caps_out,
                         uint64_t **layer_leaves_out, uint64_t **layer_digests_out,
                         uint64_t *final_poly_out, uint64_t *betas_out);
/* smallest witness w such that Poseidon(duplex state with w appended)[7] has >= pow_bits leading
 * zero bits (the reference picks any valid w via rayon find_any — SURVEY §0.5; we define the
 * minimum).  Advances the challenger like fri_proof_of_work (observe w, squeeze response). */
uint64_t fri_proof_of_work(p2o_challenger *ch, unsigned pow_bits);
/* check used on stored proofs: does w satisfy the PoW for this challenger state? (does not advance) */
int fri_pow_check(const p2o_challenger *ch, uint64_t w, unsigned pow_bits);

/* arity-2^arity_bits fold of one coset (verifier's compute_evaluation): evals[arity] ext in the
 * bit-reversed order they are stored in a layer leaf, x_index_within_coset, returns P(beta). */
void fri_compute_evaluation(uint64_t x /* base-field point of evals[x_index_within_coset] */,
                            unsigned x_index_within_coset, unsigned arity_bits, const uint64_t *evals,
                            const uint64_t beta[2], uint64_t out[2]);

/* ---- PLONK stages between the commitments (plonk.c) ---- */
enum {
  P2O_GATE_NOOP = 0,
  P2O_GATE_CONSTANT = 1,        /* p0 = num_consts */
  P2O_GATE_PUBLIC_INPUT = 2,
  P2O_GATE_ARITHMETIC = 3,      /* p0 = num_ops */
  P2O_GATE_POSEIDON = 4,
  P2O_GATE_BASE_SUM = 5,        /* BaseSumGate<2>, p0 = num_limbs */
  P2O_GATE_U32_ARITHMETIC = 6,  /* p0 = num_ops */
  P2O_GATE_U32_ADD_MANY = 7,    /* p0 = num_addends, p1 = num_ops */
  P2O_GATE_U32_SUBTRACTION = 8, /* p0 = num_ops */
  P2O_GATE_U32_RANGE_CHECK = 9, /* p0 = num_input_limbs */
  P2O_GATE_U32_INTERLEAVE = 10, /* p0 = num_ops */
  P2O_GATE_UNINTERLEAVE_TO_U32 = 11, /* p0 = num_ops */
  P2O_GATE_UNINTERLEAVE_TO_B32 = 12, /* p0 = num_ops */
  P2O_GATE_COMPARISON = 13,     /* p0 = num_bits, p1 = num_chunks */
  P2O_GATE_ARITHMETIC_EXT = 14, /* p0 = num_ops */
  P2O_GATE_MUL_EXT = 15,        /* p0 = num_ops */
  P2O_GATE_REDUCING = 16,       /* p0 = num_coeffs */
  P2O_GATE_REDUCING_EXT = 17,   /* p0 = num_coeffs */
  P2O_GATE_RANDOM_ACCESS = 18,  /* p0 = bits, p1 = num_copies | num_extra_constants << 16 */
  P2O_GATE_POSEIDON_MDS = 19,
  P2O_GATE_COSET_INTERPOLATION = 20 /* p0 = subgroup_bits, p1 = degree */
};
typedef struct {
  uint32_t kind, p0, p1;
  uint32_t selector_index;         /* selectors_info.selector_indices[row] */
  uint32_t group_start, group_end; /* selectors_info.groups[selector_index] */
  uint32_t row;                    /* index of the gate in common_data.gates */
} p2o_gate;
typedef struct {
  uint32_t degree_bits, num_wires, num_routed_wires;
  uint32_t num_constants; /* constant columns, selectors first */
  uint32_t num_selectors, num_challenges, quotient_degree_factor, num_partial_products, num_gate_constraints;
  uint32_t n_gates;
  const p2o_gate *gates;
  const uint64_t *k_is; /* num_routed_wires */
} p2o_circuit;
/* unfiltered constraints of one gate at one point: wires / consts (after the selectors) / pi_hash canonical */
unsigned plonk_eval_gate(const p2o_gate *g, const uint64_t *wires, const uint64_t *consts, const uint64_t *pi_hash,
                         uint64_t *out);
/* wires: num_wires x n, sigmas: num_routed_wires x n (values on H, column-major).
 * out: num_challenges * (1 + num_partial_products) columns x n: [Z_0.., partial products of challenge 0, ...] */
void plonk_partial_products_and_zs(const p2o_circuit *c, const uint64_t *wires, const uint64_t *sigmas,
                                   const uint64_t *betas, const uint64_t *gammas, uint64_t *out);
/* *_leaves: the three batches' LDE leaves, row-major in leaf order (as batch_from_* write them).
 * out_chunks: num_challenges x (quotient_degree_factor * n) coefficients = the quotient chunks in commit order */
void plonk_compute_quotient_polys(const p2o_circuit *c, unsigned rate_bits, const uint64_t *cs_leaves,
                                  const uint64_t *wires_leaves, const uint64_t *zs_leaves, const uint64_t *pi_hash,
                                  const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas,
                                  uint64_t *out_chunks);

/* ---- the whole prover (prove.c): what `circuit_data.prove(pw)` runs after witness generation ---- */
typedef struct {
  uint32_t rate_bits, cap_height, proof_of_work_bits, num_query_rounds;
  uint32_t n_layers;
  uint32_t reduction_arity_bits[16];
} p2o_fri_params; /* FriParams (city_common_circuit/src/verify_template/ser_data.rs:56-154); same layout as p2b_fri_params */
/* prover_data.constants_sigmas_commitment + prover_data.sigmas: built once per circuit (CircuitBuilder::build), not per proof.
 * cs_cols: num_constants + num_routed_wires columns of 2^degree_bits values on H */
typedef struct p2o_prover_data p2o_prover_data;
p2o_prover_data *p2o_prover_data_new(const p2o_circuit *c, const uint64_t *const *cs_cols, unsigned rate_bits,
                                     unsigned cap_height);
void p2o_prover_data_free(p2o_prover_data *pd);
void p2o_prover_data_cap(const p2o_prover_data *pd, uint64_t *out);
/* number of u64 words of a proof (0 on inconsistent parameters); equals p2b_proof_len */
size_t p2o_proof_len(const p2o_circuit *c, unsigned cs_cap_height, const p2o_fri_params *fp, size_t n_pis);
/* prove_with_partition_witness from the filled witness onwards; writes the words p2b_prove writes (include/p2b.h) and
 * returns their number (0 on failure).  The proof-of-work witness is the minimal one. */
size_t p2o_prove(const p2o_circuit *c, const p2o_prover_data *pd, const uint64_t *circuit_digest,
                 const uint64_t *const *wire_cols, const uint64_t *public_inputs, size_t n_pis,
                 const p2o_fri_params *fp, uint64_t *out, size_t out_cap);

unsigned p2o_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
