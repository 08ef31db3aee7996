/* p2b.h — C ABI of the B200-native Plonky2 proving hot path behind City Rollup worker jobs.
 *
 * This is the drop-in boundary.  The reference reaches the path only through
 * `circuit_data.prove(pw)` (e.g. city_common_circuit/src/proof_minifier/pm_core.rs:151,
 * city_rollup_circuit/src/block_circuits/ops/l2_transfer/circuit.rs:234,
 * city_common_circuit/src/treeprover/aggregation/state_transition/mod.rs:298); the functions replaced
 * live in the plonky2 0.2.2 dependency it patches in at Cargo.toml:101-102,128-129.  A patched
 * `plonky2` crate binds these entry points from a thin FFI crate (rust/plonky2_b200_sys, see
 * INTEGRATION.md); each declaration below names the Rust item it stands in for.
 *
 * Conventions
 *  - Field elements are raw little-endian u64 (GoldilocksField is a transparent u64).  Inputs may be
 *    non-canonical (>= p = 2^64-2^32+1, cf. city_crypto/src/hash/qhashout.rs:149-152); every output
 *    is canonical.  A digest (HashOut) is 4 u64; a cap of height h is 2^h digests; an extension
 *    element (QuadraticExtension, X^2 = 7) is 2 u64 {c0, c1}.
 *  - Every function returns P2B_OK (0) or a negative p2b_status and never throws / aborts / unwinds;
 *    the message is available from p2b_last_error().  CUDA errors are sticky: after
 *    P2B_ERR_CUDA the context is poisoned and must be destroyed (the worker loop maps this to
 *    anyhow::Error, city_rollup_core_worker/src/actors/simple.rs:83).
 *  - The caller owns every pointer it passes and every plain output buffer.  The library owns the
 *    opaque handles until the matching *_free; no handle may outlive its context.
 *  - A p2b_ctx is bound to one device and one CUDA stream and is NOT thread-safe; use one context per
 *    OS thread (several per GPU are fine and let small proofs overlap).
 *  - Host inputs may be pageable or pinned; pinned (p2b_host_alloc) buffers are copied by DMA without staging.
 *    Every entry point that takes host input returns only after that input has been read (staged or uploaded): the
 *    caller may refill its buffers at once, while the transforms are still running on the device.
 *  - There is no CPU fallback: if no CUDA device is usable p2b_init fails.
 */
#ifndef P2B_H
#define P2B_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  P2B_OK = 0,
  P2B_ERR_INVALID = -1, /* bad argument */
  P2B_ERR_CUDA = -2,    /* CUDA runtime error (context poisoned) */
  P2B_ERR_OOM = -3,     /* device or host allocation failed */
  P2B_ERR_UNSUPPORTED = -4
} p2b_status;

typedef struct p2b_ctx p2b_ctx;
typedef struct p2b_batch p2b_batch;           /* plonky2::fri::oracle::PolynomialBatch */
typedef struct p2b_tree p2b_tree;             /* plonky2::hash::merkle_tree::MerkleTree<F, PoseidonHash> */
typedef struct p2b_challenger p2b_challenger; /* plonky2::iop::challenger::Challenger<F, PoseidonHash> */

/* ---------------------------------------------------------------- context ---------------- */
int p2b_version(void);
/* Creates a context on `device` with its own non-blocking stream. */
int p2b_init(int device, p2b_ctx **out);
/* Same, but all work is enqueued on the caller's cudaStream_t (e.g. a framework's current stream);
 * the stream is borrowed, not destroyed. */
int p2b_init_on_stream(int device, void *cuda_stream, p2b_ctx **out);
void p2b_destroy(p2b_ctx *ctx);
/* Message of the last failure on this context ("" if none).  ctx may be NULL (global init error). */
const char *p2b_last_error(const p2b_ctx *ctx);
/* Blocks until all work enqueued on the context's stream has finished. */
int p2b_synchronize(p2b_ctx *ctx);
/* How the calling thread waits for the device inside the library: 0 (default) spins — lowest latency, the right
 * choice for the reference's model of one worker process per GPU (city_rollup_core_worker/src/lib.rs:131-145); 1
 * sleeps on a blocking-sync event, for hosts that run more proving threads than they have cores (several contexts
 * per GPU times several GPUs).  The environment variable P2B_SYNC=block|spin sets the default of new contexts. */
int p2b_set_blocking_sync(p2b_ctx *ctx, int on);
/* Latency mode: for a worker that has the GPU to itself (one job at a time, no sibling worker processes on the device —
 * city_rollup_core_worker/src/actors/simple.rs:32-56 run as a single process per GPU).  Dependent launch chains are
 * shortened at the price of extra work (Merkle trees fused from 2^15 digests, proof-of-work search on every SM): 3.71 ->
 * 3.35 ms per 2^12-row proof on one context; with many proofs in flight the default (throughput mode) is ~10 % faster
 * per GPU.  Same results in both modes.  Also P2B_MODE=latency in the environment.  Drops the context's prove plans. */
int p2b_set_latency_mode(p2b_ctx *ctx, int on);
/* Pinned host memory for inputs/outputs that should move by DMA without staging. */
int p2b_host_alloc(p2b_ctx *ctx, size_t bytes, void **out);
int p2b_host_free(p2b_ctx *ctx, void *p);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t p2b_launch_count(const p2b_ctx *ctx);
/* CUDA-event timing on the context's stream (bench.py times kernels on the launching stream). */
int p2b_timer_start(p2b_ctx *ctx);
int p2b_timer_stop_ms(p2b_ctx *ctx, float *ms_out); /* synchronises on the stop event */
/* Several contexts of one device working side by side (one worker thread each): ms from `first`'s
 * p2b_timer_start event to `last`'s p2b_timer_stop_ms event. */
int p2b_timer_span_ms(p2b_ctx *first, p2b_ctx *last, float *ms_out);

/* Per-stage device timing (CUDA events recorded on the context's stream around each stage of the
 * batch / FRI pipelines).  Stages: 0 h2d, 1 intt, 2 lde, 3 leaf_hash, 4 tree_levels, 5 fri_fold_ntt,
 * 6 transcript, 7 other.  p2b_profile_read synchronises, ADDS the elapsed ms of every recorded stage
 * interval since the last read to ms_out[8] and to count_out[8] (kernel launches per stage; may be NULL),
 * and clears the record. */
#define P2B_N_STAGES 8
int p2b_profile_enable(p2b_ctx *ctx, int on);
int p2b_profile_read(p2b_ctx *ctx, float *ms_out, uint64_t *count_out);

/* ---------------------------------------------------------------- PolynomialBatch -------- */
/* PolynomialBatch::from_values(values, rate_bits, blinding=false, cap_height, timing, fft_root_table)
 * cols[c] points to 2^log_n values of column c (plonky2's Vec<PolynomialValues<F>>: one allocation
 * per column).  Computes per column ifft -> lde(rate_bits) -> coset_fft(shift 7), the bit-reversed
 * leaf order and the Poseidon Merkle tree with 2^cap_height roots.  Everything stays in HBM.
 * flags: 0 or P2B_KEEP_VALUES (blinding/salting is used by no worker circuit, SURVEY §8(c), and is
 * rejected).  P2B_KEEP_VALUES keeps the input values on H in HBM next to the coefficients — the prover
 * needs witness.wire_values and prover_data.sigmas again for the partial products (p2b_zs_partial_products_commit). */
#define P2B_KEEP_VALUES 1u
int p2b_batch_from_values(p2b_ctx *ctx, const uint64_t *const *cols, size_t n_cols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch **out);
/* PolynomialBatch::from_coeffs(polynomials, ...): cols[c] = 2^log_n coefficients of column c. */
int p2b_batch_from_coeffs(p2b_ctx *ctx, const uint64_t *const *cols, size_t n_cols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch **out);
/* Device-resident inputs: d_cols is a device pointer to n_cols x 2^log_n u64, column-major
 * (column c at d_cols + c * 2^log_n).  Used when the previous prover stage already left its output
 * in HBM (witness upload once, Z/partial products, quotient chunks). */
int p2b_batch_from_values_dev(p2b_ctx *ctx, const uint64_t *d_cols, size_t n_cols, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch **out);
int p2b_batch_from_coeffs_dev(p2b_ctx *ctx, const uint64_t *d_cols, size_t n_cols, uint32_t log_n,
                              uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch **out);
void p2b_batch_free(p2b_batch *b);

size_t p2b_batch_n_cols(const p2b_batch *b);
uint32_t p2b_batch_degree_log(const p2b_batch *b);
uint32_t p2b_batch_rate_bits(const p2b_batch *b);
/* batch.merkle_tree (borrowed; freed with the batch) */
p2b_tree *p2b_batch_tree(p2b_batch *b);
/* batch.merkle_tree.cap -> out[4 << cap_height] */
int p2b_batch_cap(p2b_batch *b, uint64_t *out);
/* batch.polynomials[col].coeffs -> out[2^log_n] (natural order) */
int p2b_batch_coeffs(p2b_batch *b, size_t col, uint64_t *out);
/* batch.merkle_tree.leaves[leaf_index] (= MerkleTree::get) -> out[n_cols] */
int p2b_batch_leaf(p2b_batch *b, size_t leaf_index, uint64_t *out);
/* PolynomialBatch::get_lde_values(index, step) -> out[n_cols] (row bitrev(index*step)) */
int p2b_batch_lde_values(p2b_batch *b, size_t index, size_t step, uint64_t *out);
/* all leaves, row-major (2^(log_n+rate_bits) x n_cols), i.e. batch.merkle_tree.leaves flattened */
int p2b_batch_leaves(p2b_batch *b, uint64_t *out);
/* Device views (valid until p2b_batch_free): LDE values column-major in leaf order
 * (column c, leaf j at d_lde[c * 2^(log_n+rate_bits) + j]) and coefficients column-major. */
const uint64_t *p2b_batch_dev_lde(const p2b_batch *b);
const uint64_t *p2b_batch_dev_coeffs(const p2b_batch *b);

/* the whole LDE of polynomial `col` in leaf order: out[j] = batch.merkle_tree.leaves[j][col], j < 2^(log_n+rate_bits) */
int p2b_batch_lde_col(p2b_batch *b, size_t col, uint64_t *out);
/* values on H of column `col` (batches built from values with P2B_KEEP_VALUES) -> out[2^log_n] */
int p2b_batch_values(p2b_batch *b, size_t col, uint64_t *out);

/* ---- circuit-data reuse.  prover_data.constants_sigmas_commitment is built once per circuit
 * (CircuitBuilder::build: city_rollup_core_worker_qbench/src/qbench.rs:21, city_rollup_circuit/src/sighash_circuits/
 * sighash_wrapper.rs:142-148) and read by every proof of it. */
/* A read-only view of `src` for another context of the SAME device: the contexts of a GPU then share one device copy of
 * the constants|sigmas batch.  `src` (and its context) must outlive the view; freeing the view frees nothing else. */
int p2b_batch_attach(p2b_ctx *ctx, const p2b_batch *src, p2b_batch **out);
/* Serialised form (host bytes): header, cap, coefficients (+ the values on H of a P2B_KEEP_VALUES batch).  import
 * recomputes the LDE and the Merkle tree on the device and fails with P2B_ERR_INVALID when the recomputed cap differs
 * from the stored one. */
size_t p2b_batch_export_len(const p2b_batch *b);
int p2b_batch_export(p2b_batch *b, uint8_t *out, size_t out_cap, size_t *written);
int p2b_batch_import(p2b_ctx *ctx, const uint8_t *bytes, size_t n_bytes, p2b_batch **out);

/* ---------------------------------------------------------------- PLONK stages ----------- */
/* What the prover stages between the commitments need of plonky2's CommonCircuitData
 * (the reference dumps one at city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145):
 * the gate list with its selector layout (common_data.gates / selectors_info), the wire and constant
 * counts, num_challenges, quotient_degree_factor, num_partial_products, k_is.  Lookups and blinding are
 * not supported (no worker circuit uses them: num_lookup_polys 0, zero_knowledge false). */
typedef enum {
  P2B_GATE_NOOP = 0,            /* plonky2 NoopGate */
  P2B_GATE_CONSTANT = 1,        /* ConstantGate { num_consts = p0 } */
  P2B_GATE_PUBLIC_INPUT = 2,    /* PublicInputGate */
  P2B_GATE_ARITHMETIC = 3,      /* ArithmeticGate { num_ops = p0 } */
  P2B_GATE_POSEIDON = 4,        /* PoseidonGate */
  P2B_GATE_BASE_SUM = 5,        /* BaseSumGate<2> { num_limbs = p0 } */
  P2B_GATE_U32_ARITHMETIC = 6,  /* city_common_circuit/src/u32/gates/arithmetic_u32.rs  { num_ops = p0 } */
  P2B_GATE_U32_ADD_MANY = 7,    /* .../add_many_u32.rs { num_addends = p0, num_ops = p1 } */
  P2B_GATE_U32_SUBTRACTION = 8, /* .../subtraction_u32.rs { num_ops = p0 } */
  P2B_GATE_U32_RANGE_CHECK = 9, /* .../range_check_u32.rs { num_input_limbs = p0 } */
  P2B_GATE_U32_INTERLEAVE = 10, /* .../interleave_u32.rs { num_ops = p0 } */
  P2B_GATE_UNINTERLEAVE_TO_U32 = 11, /* .../uninterleave_to_u32.rs { num_ops = p0 } */
  P2B_GATE_UNINTERLEAVE_TO_B32 = 12, /* .../uninterleave_to_b32.rs { num_ops = p0 } */
  P2B_GATE_COMPARISON = 13,     /* .../comparison.rs { num_bits = p0, num_chunks = p1 } ((32, 16) in pad_circuit.rs:33) */
  P2B_GATE_ARITHMETIC_EXT = 14, /* plonky2 ArithmeticExtensionGate { num_ops = p0 } */
  P2B_GATE_MUL_EXT = 15,        /* MulExtensionGate { num_ops = p0 } */
  P2B_GATE_REDUCING = 16,       /* ReducingGate { num_coeffs = p0 } (43 in pad_circuit.rs) */
  P2B_GATE_REDUCING_EXT = 17,   /* ReducingExtensionGate { num_coeffs = p0 } (32) */
  P2B_GATE_RANDOM_ACCESS = 18,  /* RandomAccessGate { bits = p0, num_copies = p1 & 0xFFFF, num_extra_constants = p1 >> 16 } */
  P2B_GATE_POSEIDON_MDS = 19,   /* PoseidonMdsGate */
  P2B_GATE_COSET_INTERPOLATION = 20 /* CosetInterpolationGate { subgroup_bits = p0, degree = p1 } (with_max_degree(4, 8): degree 6) */
} p2b_gate_kind;
typedef struct {
  uint32_t kind, p0, p1;
  uint32_t selector_index;         /* selectors_info.selector_indices[row]: constants column of its selector */
  uint32_t group_start, group_end; /* selectors_info.groups[selector_index] */
  uint32_t row;                    /* index of the gate in common_data.gates */
} p2b_gate;
typedef struct {
  uint32_t degree_bits, num_wires, num_routed_wires;
  uint32_t num_constants; /* constant columns of constants_sigmas, selectors first */
  uint32_t num_selectors, num_challenges, quotient_degree_factor, num_partial_products, num_gate_constraints;
  uint32_t n_gates;
  const p2b_gate *gates;
  const uint64_t *k_is; /* num_routed_wires */
} p2b_circuit_desc;
typedef struct p2b_circuit p2b_circuit;
int p2b_circuit_new(p2b_ctx *ctx, const p2b_circuit_desc *desc, p2b_circuit **out);
void p2b_circuit_free(p2b_circuit *c);

/* plonk::prover::all_wires_permutation_partial_products (wires_permutation_partial_products_and_zs for every
 * challenge) followed by the PolynomialBatch::from_values of [Z_0.., partial products of challenge 0, ...]
 * exactly as prove_with_partition_witness orders them.  constants_sigmas and wires must have been built
 * from values with P2B_KEEP_VALUES.  betas / gammas: num_challenges elements each.  The result keeps its
 * values too (p2b_batch_values). */
int p2b_zs_partial_products_commit(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
                                   const p2b_batch *wires, const uint64_t *betas, const uint64_t *gammas,
                                   uint32_t rate_bits, uint32_t cap_height, p2b_batch **out);
/* plonk::prover::compute_quotient_polys + the chunking + PolynomialBatch::from_coeffs of the
 * num_challenges * quotient_degree_factor quotient chunks.  quotient_degree_factor must be a power of two
 * <= 2^rate_bits.  pi_hash = public_inputs_hash (4 elements). */
int p2b_quotient_commit(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
                        const p2b_batch *wires, const p2b_batch *zs_partial_products, const uint64_t *pi_hash,
                        const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, uint32_t rate_bits,
                        uint32_t cap_height, p2b_batch **out);

/* ---------------------------------------------------------------- MerkleTree ------------- */
/* MerkleTree::<F, PoseidonHash>::new(leaves, cap_height); leaves row-major n_leaves x leaf_len
 * (n_leaves a power of two, cap_height <= log2(n_leaves)). */
int p2b_merkle_new(p2b_ctx *ctx, const uint64_t *leaves, size_t n_leaves, size_t leaf_len,
                   uint32_t cap_height, p2b_tree **out);
void p2b_tree_free(p2b_tree *t);
size_t p2b_tree_n_leaves(const p2b_tree *t);
uint32_t p2b_tree_cap_height(const p2b_tree *t);
/* tree.cap -> out[4 << cap_height] */
int p2b_tree_cap(p2b_tree *t, uint64_t *out);
/* MerkleTree::prove(leaf_index).siblings -> out[4 * (log2(n_leaves) - cap_height)], leaf level first */
int p2b_tree_prove(p2b_tree *t, size_t leaf_index, uint64_t *out);
/* tree.digests in plonky2's interleaved layout -> out[4 * 2 * (n_leaves - 2^cap_height)] */
int p2b_tree_digests(p2b_tree *t, uint64_t *out);
/* tree.leaves[leaf_index] -> out[leaf_len] (trees built by p2b_merkle_new / p2b_fri_commit) */
int p2b_tree_leaf(p2b_tree *t, size_t leaf_index, uint64_t *out);

/* ---------------------------------------------------------------- Poseidon --------------- */
/* n independent permutations; states is n x 12 (row-major), permuted in place.
 * (plonky2::hash::poseidon::Poseidon::poseidon) */
int p2b_poseidon_permute(p2b_ctx *ctx, uint64_t *states, size_t n);
/* PoseidonHash::hash_no_pad(in[0..len]) -> out[4] */
int p2b_hash_no_pad(p2b_ctx *ctx, const uint64_t *in, size_t len, uint64_t *out);
/* PoseidonHash::two_to_one for n pairs: left/right are n x 4, out n x 4 */
int p2b_two_to_one(p2b_ctx *ctx, const uint64_t *left, const uint64_t *right, size_t n, uint64_t *out);

/* ---------------------------------------------------------------- Challenger ------------- */
/* The transcript lives in HBM so that the FRI commit loop (tree -> observe cap -> beta -> fold) runs
 * without a host round trip per layer. */
int p2b_challenger_new(p2b_ctx *ctx, p2b_challenger **out);
void p2b_challenger_free(p2b_challenger *c);
/* Challenger::observe_elements */
int p2b_challenger_observe(p2b_challenger *c, const uint64_t *elems, size_t n);
/* Challenger::observe_cap(&tree.cap) without leaving the device */
int p2b_challenger_observe_cap(p2b_challenger *c, p2b_tree *t);
/* Challenger::get_n_challenges(n) -> out[n] (synchronises) */
int p2b_challenger_get(p2b_challenger *c, size_t n, uint64_t *out);
/* raw state for seeding / cross-checking: out = sponge_state[12] || input_len || input_buffer[8] ||
 * output_len || output_buffer[8]  (30 u64) */
int p2b_challenger_export(p2b_challenger *c, uint64_t *out30);
int p2b_challenger_import(p2b_challenger *c, const uint64_t *in30);

/* ---------------------------------------------------------------- FRI -------------------- */
/* fri::prover::fri_committed_trees(coeffs, values, challenger, fri_params):
 * coeffs / values: `len` extension elements (interleaved c0,c1), values in natural order on the coset
 * 7*<w_len>.  For each layer i (arity 2^arity_bits[i]): Merkle tree over the bit-reversed values
 * chunked by arity, observe its cap, squeeze beta, fold the coefficients, coset-NTT the next layer.
 * layers_out[i] receives the layer tree (caller frees with p2b_tree_free).  final_poly_out receives
 * (len >> sum(arity_bits) >> rate_bits) extension elements, which are also observed. */
int p2b_fri_commit(p2b_ctx *ctx, const uint64_t *coeffs_ext, const uint64_t *values_ext, size_t len,
                   const uint32_t *arity_bits, size_t n_layers, uint32_t rate_bits, uint32_t cap_height,
                   p2b_challenger *challenger, p2b_tree **layers_out, uint64_t *final_poly_out);
/* fri::prover::fri_proof_of_work: the SMALLEST witness w with >= pow_bits leading zero bits in the
 * response (the reference accepts any valid w, found by rayon find_any — SURVEY §0.5).  Observes w and
 * squeezes the response like the reference. */
int p2b_fri_pow(p2b_ctx *ctx, p2b_challenger *challenger, uint32_t pow_bits, uint64_t *witness_out);

/* ---------------------------------------------------------------- openings + FRI proof --- */
/* OpeningSet::new's eval_commitment: polynomials first .. first+count of the batch, evaluated at the
 * extension point {point[0], point[1]} -> out[2 * count] */
int p2b_batch_eval_ext(p2b_batch *b, const uint64_t *point, size_t first, size_t count, uint64_t *out);

/* FriBatchInfo: an opening point and the polynomials opened there, as ranges of (oracle, first, count) */
#define P2B_MAX_FRI_RANGES 8
typedef struct {
  uint64_t point[2];
  uint32_t n_ranges;
  struct {
    uint32_t oracle, first, count;
  } ranges[P2B_MAX_FRI_RANGES];
} p2b_fri_batch;
/* FriParams / FriConfig (city_common_circuit/src/verify_template/ser_data.rs:56-154) */
#define P2B_MAX_FRI_LAYERS 16
typedef struct {
  uint32_t rate_bits, cap_height, proof_of_work_bits, num_query_rounds;
  uint32_t n_layers;
  uint32_t reduction_arity_bits[P2B_MAX_FRI_LAYERS];
} p2b_fri_params;
/* Number of u64 words p2b_prove_openings writes for these oracles and parameters. */
size_t p2b_fri_proof_len(const p2b_batch *const *oracles, size_t n_oracles, const p2b_fri_params *params);
/* PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing): squeezes alpha, builds
 * final_poly = sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i) over the batches, its LDE, then fri_proof:
 * commit phase, proof of work (minimal witness), query rounds.  All oracles must share degree and rate_bits =
 * params->rate_bits.  Output (u64 words, no length prefixes, every element canonical) in FriProof's field order:
 *   commit_phase_merkle_caps   n_layers x (4 << cap_height)
 *   query_round_proofs         num_query_rounds x { per oracle: leaf (n_cols) , siblings 4 x (log2(leaves) - cap_height);
 *                                                   per layer: evals 2 x arity, siblings 4 x (layer height - cap_height) }
 *   final_poly                 2 x (n << rate_bits >> sum(arity_bits) >> rate_bits)
 *   pow_witness                1 */
int p2b_prove_openings(p2b_ctx *ctx, const p2b_batch *const *oracles, size_t n_oracles, const p2b_fri_batch *batches,
                       size_t n_batches, p2b_challenger *challenger, const p2b_fri_params *params, uint64_t *proof_out,
                       size_t proof_cap);

/* ---------------------------------------------------------------- prove ------------------ */
/* plonk::prover::prove_with_partition_witness from the filled witness onwards — what every
 * `circuit_data.prove(pw)` call site of the reference spends its time in after witness
 * generation — in ONE call: wires commit, betas/gammas, Z / partial products commit, alphas, quotient commit,
 * zeta, openings, prove_openings.  The transcript stays on the device: the only host synchronisations are the
 * proof-of-work search and the final download.  wire_cols[w] = witness.wire_values[w] (2^degree_bits values);
 * constants_sigmas = prover_data.constants_sigmas_commitment (built with P2B_KEEP_VALUES).
 * Output (u64 words, canonical, no length prefixes) in ProofWithPublicInputs' field order:
 *   wires_cap | plonk_zs_partial_products_cap | quotient_polys_cap              3 x (4 << cap_height)
 *   openings: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys
 *             (2 words per extension element)
 *   opening_proof: as p2b_prove_openings writes it
 *   public_inputs                                                               n_public_inputs
 * The proof-of-work witness is the minimal one (the reference's is schedule dependent). */
size_t p2b_proof_len(const p2b_circuit *circuit, const p2b_batch *constants_sigmas, const p2b_fri_params *params,
                     size_t n_public_inputs);
int p2b_prove(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
              const uint64_t *circuit_digest, const uint64_t *const *wire_cols, const uint64_t *public_inputs,
              size_t n_public_inputs, const p2b_fri_params *params, uint64_t *proof_out, size_t proof_cap);

/* The same with the witness already in HBM: d_wire_values is a device pointer to num_wires x 2^degree_bits u64,
 * column-major (a GPU witness generator, or bench.py's device-resident arm).  The proof still lands on the host. */
int p2b_prove_dev(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
                  const uint64_t *circuit_digest, const uint64_t *d_wire_values, const uint64_t *public_inputs,
                  size_t n_public_inputs, const p2b_fri_params *params, uint64_t *proof_out, size_t proof_cap);

/* Asynchronous form for a worker that generates the next witness while the GPU proves (the reference's worker does
 * witness generation and then a blocking prove, city_rollup_circuit/src/worker/traits.rs:143-160; the sighash jobs even
 * prove a STARK inside witness generation, city_common_circuit/src/hash/accelerator/sha256/smartgadget.rs:518-523).
 * p2b_prove_submit enqueues the whole proof and returns: the witness has been staged (pageable columns) or its upload
 * has completed (pinned columns), so every buffer passed in may be reused at once.  One proof may be pending per
 * context; one host thread can keep several contexts busy.  p2b_prove_poll: 1 = finished, 0 = still running.
 * p2b_prove_collect waits and writes the proof words (as p2b_prove). */
int p2b_prove_submit(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
                     const uint64_t *circuit_digest, const uint64_t *const *wire_cols, const uint64_t *public_inputs,
                     size_t n_public_inputs, const p2b_fri_params *params);
int p2b_prove_poll(p2b_ctx *ctx);
int p2b_prove_collect(p2b_ctx *ctx, uint64_t *proof_out, size_t proof_cap);
/* p2b_prove_submit without the wait for the upload of pinned witness columns: they must stay untouched until
 * p2b_prove_upload_poll returns 1 (or the proof has been collected).  For a thread that drives many contexts: with N
 * proofs in flight an upload can wait milliseconds behind other contexts' kernels in a shared hardware queue, and a
 * wait inside submit would stall every other context of that thread (measured with 24 contexts on one thread: 143
 * proofs/s with the wait).  p2b_prove_upload_poll: 1 = the witness has been read, 0 = not yet; when a prove plan is
 * replayed from a contiguous pinned matrix the upload is a node of the captured graph reading the caller's memory
 * directly (no host copy inside submit), and the poll turns 1 when the proof has finished. */
int p2b_prove_submit_nowait(p2b_ctx *ctx, const p2b_circuit *circuit, const p2b_batch *constants_sigmas,
                            const uint64_t *circuit_digest, const uint64_t *const *wire_cols,
                            const uint64_t *public_inputs, size_t n_public_inputs, const p2b_fri_params *params);
int p2b_prove_upload_poll(p2b_ctx *ctx);
/* Prove plans: from the second proof of a given (circuit, constants_sigmas, parameters) on, a context replays the proof
 * as ONE captured CUDA graph (P2B_GRAPH=0 turns this off).  Counts of this context's plans by state and the number of
 * kernels one replay launches; after a failed capture p2b_last_error tells why that shape stays on the eager path. */
int p2b_plan_info(p2b_ctx *ctx, uint32_t *n_ready, uint32_t *n_seen, uint32_t *n_failed, uint32_t *kernels_per_launch);

/* ---------------------------------------------------------------- proof bytes ------------ */
/* The byte form the reference stores and ships proofs in: `bincode::serialize(&ProofWithPublicInputs)`
 * (city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46,65-72; city_redis_store/src/lib.rs:54-82) —
 * bincode 1.3.3 defaults: little endian, every Vec prefixed by its u64 length.  The words p2b_prove writes are the
 * same fields in the same order without the prefixes; these host-only helpers add / strip them for a given circuit
 * shape (CommonCircuitData: city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145).  Pinned on the ten
 * stored proofs of qbench_data/example.bin (tests/test_proof_bincode.py): blob -> words -> identical blob. */
typedef struct {
  uint32_t degree_bits;
  uint32_t num_constants, num_routed_wires, num_wires; /* constants|sigmas width = num_constants + num_routed_wires */
  uint32_t num_challenges, num_partial_products, quotient_degree_factor;
  uint32_t constants_sigmas_cap_height; /* cap height of the circuit's constants_sigmas tree (= fri cap_height in plonky2) */
  uint32_t n_public_inputs;
} p2b_proof_shape;
/* number of u64 words of a proof of this shape (= p2b_proof_len for the matching circuit); 0 on invalid arguments */
size_t p2b_proof_words(const p2b_proof_shape *shape, const p2b_fri_params *params);
/* number of bytes of its bincode form; 0 on invalid arguments */
size_t p2b_proof_bincode_len(const p2b_proof_shape *shape, const p2b_fri_params *params);
/* words (as written by p2b_prove) -> bincode bytes; *written (optional) = bytes produced */
int p2b_proof_to_bincode(const p2b_proof_shape *shape, const p2b_fri_params *params, const uint64_t *words,
                         size_t n_words, uint8_t *out, size_t out_cap, size_t *written);
/* bincode bytes -> words; every length prefix is checked against the shape (P2B_ERR_INVALID on any mismatch, so a
 * proof of another circuit is rejected instead of mis-sliced) */
int p2b_proof_from_bincode(const p2b_proof_shape *shape, const p2b_fri_params *params, const uint8_t *bytes,
                           size_t n_bytes, uint64_t *words_out, size_t words_cap, size_t *n_words);

#ifdef __cplusplus
}
#endif
#endif /* P2B_H */
