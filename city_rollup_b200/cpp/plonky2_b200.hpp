// plonky2_b200.hpp — header-only C++17 host mirror of the plonky2 prover surface City Rollup's workers
// call (PolynomialBatch::from_values / from_coeffs, MerkleTree::new / prove, Challenger,
// fri_committed_trees, fri_proof_of_work), over the C ABI of include/p2b.h.  RAII handles, errors as
// exceptions (the analogue of the anyhow::Result the reference propagates,
// city_rollup_core_worker/src/actors/simple.rs:83).  No computation happens on the host.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/p2b.h"

namespace plonky2_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("p2b error " + std::to_string(c) + ": " + m), code(c) {}
};

using F = uint64_t;                 // GoldilocksField (raw u64)
using HashOut = std::array<F, 4>;   // plonky2::hash::hash_types::HashOut
using Ext = std::array<F, 2>;       // QuadraticExtension<GoldilocksField>

class Context {
 public:
  explicit Context(int device = 0) {
    int rc = p2b_init(device, &h_);
    if (rc != P2B_OK) throw Error(rc, p2b_last_error(nullptr));
  }
  ~Context() { p2b_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  void check(int rc) const {
    if (rc != P2B_OK) throw Error(rc, p2b_last_error(h_));
  }
  p2b_ctx* get() const { return h_; }
  // sleep instead of spinning while waiting for the device (more proving threads than host cores)
  void set_blocking_sync(bool on) const { check(p2b_set_blocking_sync(h_, on ? 1 : 0)); }
  // one proof at a time on this GPU: shorter launch chains for more work (p2b.h); the default is throughput mode
  void set_latency_mode(bool on) const { check(p2b_set_latency_mode(h_, on ? 1 : 0)); }

 private:
  p2b_ctx* h_ = nullptr;
};

// plonky2::hash::merkle_tree::MerkleTree<F, PoseidonHash>
class MerkleTree {
 public:
  // MerkleTree::new(leaves, cap_height); leaves row-major n_leaves x leaf_len
  static MerkleTree create(const Context& ctx, const std::vector<F>& leaves, size_t n_leaves, size_t leaf_len, uint32_t cap_height) {
    p2b_tree* t = nullptr;
    ctx.check(p2b_merkle_new(ctx.get(), leaves.data(), n_leaves, leaf_len, cap_height, &t));
    return MerkleTree(&ctx, t, true);
  }
  MerkleTree(const Context* ctx, p2b_tree* t, bool owned) : ctx_(ctx), t_(t), owned_(owned) {}
  MerkleTree(MerkleTree&& o) noexcept : ctx_(o.ctx_), t_(std::exchange(o.t_, nullptr)), owned_(o.owned_) {}
  ~MerkleTree() { if (owned_ && t_) p2b_tree_free(t_); }
  std::vector<HashOut> cap() const {
    std::vector<HashOut> out(size_t(1) << p2b_tree_cap_height(t_));
    ctx_->check(p2b_tree_cap(t_, out[0].data()));
    return out;
  }
  // MerkleTree::prove(leaf_index).siblings
  std::vector<HashOut> prove(size_t leaf_index) const {
    size_t n = p2b_tree_n_leaves(t_), L = 0;
    while ((size_t(1) << L) < n) L++;
    L -= p2b_tree_cap_height(t_);
    std::vector<HashOut> out(L ? L : 1);
    ctx_->check(p2b_tree_prove(t_, leaf_index, out[0].data()));
    out.resize(L);
    return out;
  }
  p2b_tree* get() const { return t_; }

 private:
  const Context* ctx_;
  p2b_tree* t_;
  bool owned_;
};

// plonky2::fri::oracle::PolynomialBatch<F, PoseidonGoldilocksConfig, 2>
class PolynomialBatch {
 public:
  // from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table)
  // keep_values keeps the values on H in HBM (the witness columns / sigmas the Z stage reads again)
  static PolynomialBatch from_values(const Context& ctx, const std::vector<std::vector<F>>& values, size_t rate_bits,
                                     bool blinding, size_t cap_height, bool keep_values = false) {
    return build(ctx, values, rate_bits, blinding, cap_height, true, keep_values);
  }
  // from_coeffs(polynomials, rate_bits, blinding, cap_height, timing, fft_root_table)
  static PolynomialBatch from_coeffs(const Context& ctx, const std::vector<std::vector<F>>& polys, size_t rate_bits,
                                     bool blinding, size_t cap_height) {
    return build(ctx, polys, rate_bits, blinding, cap_height, false, false);
  }
  // adopt a handle produced by a prover stage (Z / partial products, quotient chunks)
  PolynomialBatch(const Context& c, p2b_batch* b) : ctx_(&c), b_(b) {}
  // OpeningSet's eval_commitment: polynomials [first, first + count) at an extension point
  std::vector<Ext> eval_ext(const Ext& point, size_t first, size_t count) const {
    std::vector<Ext> out(count ? count : 1);
    ctx_->check(p2b_batch_eval_ext(b_, point.data(), first, count, out[0].data()));
    out.resize(count);
    return out;
  }
  PolynomialBatch(PolynomialBatch&& o) noexcept : ctx_(o.ctx_), b_(std::exchange(o.b_, nullptr)) {}
  ~PolynomialBatch() { if (b_) p2b_batch_free(b_); }
  MerkleTree merkle_tree() const { return MerkleTree(ctx_, p2b_batch_tree(b_), false); }
  // get_lde_values(index, step)
  std::vector<F> get_lde_values(size_t index, size_t step) const {
    std::vector<F> out(p2b_batch_n_cols(b_));
    ctx_->check(p2b_batch_lde_values(b_, index, step, out.data()));
    return out;
  }
  std::vector<F> coeffs(size_t col) const {
    std::vector<F> out(size_t(1) << p2b_batch_degree_log(b_));
    ctx_->check(p2b_batch_coeffs(b_, col, out.data()));
    return out;
  }
  p2b_batch* get() const { return b_; }
  // A read-only view for another context of the same device (p2b_batch_attach): the worker threads of a GPU share ONE
  // device copy of a circuit's constants|sigmas batch.  *this must outlive the view.
  PolynomialBatch attach(const Context& other) const {
    p2b_batch* v = nullptr;
    other.check(p2b_batch_attach(other.get(), b_, &v));
    return PolynomialBatch(&other, v);
  }
  // bytes a later process imports instead of re-running CircuitBuilder::build's commitment (p2b_batch_export / import)
  std::vector<uint8_t> export_bytes() const {
    std::vector<uint8_t> out(p2b_batch_export_len(b_));
    size_t written = 0;
    ctx_->check(p2b_batch_export(b_, out.data(), out.size(), &written));
    out.resize(written);
    return out;
  }
  static PolynomialBatch import_bytes(const Context& ctx, const std::vector<uint8_t>& bytes) {
    p2b_batch* b = nullptr;
    ctx.check(p2b_batch_import(ctx.get(), bytes.data(), bytes.size(), &b));
    return PolynomialBatch(&ctx, b);
  }

 private:
  PolynomialBatch(const Context* c, p2b_batch* b) : ctx_(c), b_(b) {}
  static PolynomialBatch build(const Context& ctx, const std::vector<std::vector<F>>& cols, size_t rate_bits, bool blinding,
                               size_t cap_height, bool values, bool keep_values) {
    if (cols.empty()) throw Error(P2B_ERR_INVALID, "empty batch");
    size_t n = cols[0].size();
    uint32_t log_n = 0;
    while ((size_t(1) << log_n) < n) log_n++;
    if ((size_t(1) << log_n) != n) throw Error(P2B_ERR_INVALID, "length must be a power of two");
    std::vector<const F*> ptrs;
    for (auto& c : cols) {
      if (c.size() != n) throw Error(P2B_ERR_INVALID, "ragged columns");
      ptrs.push_back(c.data());
    }
    p2b_batch* b = nullptr;
    auto fn = values ? p2b_batch_from_values : p2b_batch_from_coeffs;
    ctx.check(fn(ctx.get(), ptrs.data(), ptrs.size(), log_n, (uint32_t)rate_bits, (uint32_t)cap_height, (blinding ? 2u : 0u) | (keep_values ? P2B_KEEP_VALUES : 0u), &b));
    return PolynomialBatch(&ctx, b);
  }
  const Context* ctx_;
  p2b_batch* b_;
};

// plonky2::iop::challenger::Challenger<F, PoseidonHash>
class Challenger {
 public:
  explicit Challenger(const Context& ctx) : ctx_(&ctx) { ctx.check(p2b_challenger_new(ctx.get(), &c_)); }
  ~Challenger() { p2b_challenger_free(c_); }
  Challenger(const Challenger&) = delete;
  void observe_elements(const std::vector<F>& e) { ctx_->check(p2b_challenger_observe(c_, e.data(), e.size())); }
  void observe_cap(const MerkleTree& t) { ctx_->check(p2b_challenger_observe_cap(c_, t.get())); }
  std::vector<F> get_n_challenges(size_t n) {
    std::vector<F> out(n ? n : 1);
    ctx_->check(p2b_challenger_get(c_, n, out.data()));
    out.resize(n);
    return out;
  }
  Ext get_extension_challenge() {
    auto v = get_n_challenges(2);
    return {v[0], v[1]};
  }
  p2b_challenger* get() const { return c_; }

 private:
  const Context* ctx_;
  p2b_challenger* c_ = nullptr;
};

// fri::prover::fri_committed_trees(coeffs, values, challenger, fri_params)
inline std::pair<std::vector<MerkleTree>, std::vector<Ext>> fri_committed_trees(
    const Context& ctx, const std::vector<Ext>& coeffs, const std::vector<Ext>& values, Challenger& challenger,
    const std::vector<uint32_t>& reduction_arity_bits, uint32_t rate_bits, uint32_t cap_height) {
  size_t len = coeffs.size(), n_final = len >> rate_bits;
  for (uint32_t a : reduction_arity_bits) n_final >>= a;
  std::vector<p2b_tree*> raw(reduction_arity_bits.size() ? reduction_arity_bits.size() : 1, nullptr);
  std::vector<Ext> final_poly(n_final ? n_final : 1);
  ctx.check(p2b_fri_commit(ctx.get(), coeffs[0].data(), values[0].data(), len, reduction_arity_bits.data(),
                           reduction_arity_bits.size(), rate_bits, cap_height, challenger.get(), raw.data(),
                           final_poly[0].data()));
  std::vector<MerkleTree> trees;
  for (size_t i = 0; i < reduction_arity_bits.size(); i++) trees.emplace_back(&ctx, raw[i], true);
  final_poly.resize(n_final);
  return {std::move(trees), std::move(final_poly)};
}

// fri::prover::fri_proof_of_work(challenger, config) — returns the minimal witness
// get_circuit_fingerprint_generic(verifier_data) = hash_no_pad(constants_sigmas_cap || circuit_digest)
// (city_common_circuit/src/proof_minifier/pm_core.rs:18-42): the name City Rollup gives a circuit (whitelist leaves,
// allowed_fingerprints of the aggregators)
inline HashOut circuit_fingerprint(const Context& ctx, const PolynomialBatch& constants_sigmas, const HashOut& circuit_digest) {
  std::vector<F> all;
  for (const HashOut& h : constants_sigmas.merkle_tree().cap()) all.insert(all.end(), h.begin(), h.end());
  all.insert(all.end(), circuit_digest.begin(), circuit_digest.end());
  HashOut out{};
  ctx.check(p2b_hash_no_pad(ctx.get(), all.data(), all.size(), out.data()));
  return out;
}
inline F fri_proof_of_work(const Context& ctx, Challenger& challenger, uint32_t proof_of_work_bits) {
  F w = 0;
  ctx.check(p2b_fri_pow(ctx.get(), challenger.get(), proof_of_work_bits, &w));
  return w;
}

// The slice of CommonCircuitData the prover stages read (gates + selectors_info, counts, k_is), uploaded once
class CircuitData {
 public:
  CircuitData(const Context& ctx, p2b_circuit_desc desc, std::vector<p2b_gate> gates, std::vector<F> k_is)
      : ctx_(&ctx), gates_(std::move(gates)), k_is_(std::move(k_is)), desc_(desc) {
    desc_.gates = gates_.data();
    desc_.n_gates = (uint32_t)gates_.size();
    desc_.k_is = k_is_.data();
    ctx.check(p2b_circuit_new(ctx.get(), &desc_, &c_));
  }
  ~CircuitData() { p2b_circuit_free(c_); }
  CircuitData(const CircuitData&) = delete;
  const p2b_circuit_desc& desc() const { return desc_; }
  p2b_circuit* get() const { return c_; }

 private:
  const Context* ctx_;
  std::vector<p2b_gate> gates_;
  std::vector<F> k_is_;
  p2b_circuit_desc desc_;
  p2b_circuit* c_ = nullptr;
};

// plonk::prover::all_wires_permutation_partial_products + the commit of [Zs, partial products]
inline PolynomialBatch all_wires_permutation_partial_products(const Context& ctx, const CircuitData& circuit,
                                                              const PolynomialBatch& constants_sigmas, const PolynomialBatch& wires,
                                                              const std::vector<F>& betas, const std::vector<F>& gammas,
                                                              uint32_t rate_bits, uint32_t cap_height) {
  p2b_batch* out = nullptr;
  ctx.check(p2b_zs_partial_products_commit(ctx.get(), circuit.get(), constants_sigmas.get(), wires.get(), betas.data(),
                                           gammas.data(), rate_bits, cap_height, &out));
  return PolynomialBatch(ctx, out);
}

// plonk::prover::compute_quotient_polys + chunking + from_coeffs
inline PolynomialBatch compute_quotient_polys(const Context& ctx, const CircuitData& circuit, const PolynomialBatch& constants_sigmas,
                                              const HashOut& public_inputs_hash, const PolynomialBatch& wires,
                                              const PolynomialBatch& zs_partial_products, const std::vector<F>& betas,
                                              const std::vector<F>& gammas, const std::vector<F>& alphas, uint32_t rate_bits,
                                              uint32_t cap_height) {
  p2b_batch* out = nullptr;
  ctx.check(p2b_quotient_commit(ctx.get(), circuit.get(), constants_sigmas.get(), wires.get(), zs_partial_products.get(),
                                public_inputs_hash.data(), betas.data(), gammas.data(), alphas.data(), rate_bits, cap_height, &out));
  return PolynomialBatch(ctx, out);
}

// Pinned host matrix (p2b_host_alloc): witness columns written here go to the device by one DMA instead of being
// staged through the context's pinned double buffer.  Row r = column r of the batch.
class PinnedColumns {
 public:
  PinnedColumns(const Context& ctx, size_t n_cols, size_t n) : ctx_(&ctx), n_cols_(n_cols), n_(n) {
    void* p = nullptr;
    ctx.check(p2b_host_alloc(ctx.get(), n_cols * n * sizeof(F), &p));
    base_ = static_cast<F*>(p);
    for (size_t c = 0; c < n_cols; c++) ptrs_.push_back(base_ + c * n);
  }
  ~PinnedColumns() { p2b_host_free(ctx_->get(), base_); }
  PinnedColumns(const PinnedColumns&) = delete;
  F* column(size_t c) { return base_ + c * n_; }
  void fill(const std::vector<std::vector<F>>& cols) {
    for (size_t c = 0; c < n_cols_ && c < cols.size(); c++) std::copy(cols[c].begin(), cols[c].end(), column(c));
  }
  const std::vector<const F*>& pointers() const { return ptrs_; }

 private:
  const Context* ctx_;
  size_t n_cols_, n_;
  F* base_ = nullptr;
  std::vector<const F*> ptrs_;
};

// plonk::prover::prove_with_partition_witness after witness generation, one library call; returns the flat proof words
// (ProofWithPublicInputs field order, include/p2b.h).  wire_cols: one pointer per witness column.
inline std::vector<F> prove(const Context& ctx, const CircuitData& circuit, const PolynomialBatch& constants_sigmas,
                            const HashOut& circuit_digest, const std::vector<const F*>& wire_cols,
                            const std::vector<F>& public_inputs, const p2b_fri_params& params) {
  std::vector<F> out(p2b_proof_len(circuit.get(), constants_sigmas.get(), &params, public_inputs.size()));
  if (out.empty()) throw Error(P2B_ERR_INVALID, "inconsistent FRI parameters");
  if (wire_cols.size() != circuit.desc().num_wires) throw Error(P2B_ERR_INVALID, "witness column count does not match the circuit");
  ctx.check(p2b_prove(ctx.get(), circuit.get(), constants_sigmas.get(), circuit_digest.data(), wire_cols.data(), public_inputs.data(),
                      public_inputs.size(), &params, out.data(), out.size()));
  return out;
}
inline std::vector<F> prove(const Context& ctx, const CircuitData& circuit, const PolynomialBatch& constants_sigmas,
                            const HashOut& circuit_digest, const std::vector<std::vector<F>>& wire_values,
                            const std::vector<F>& public_inputs, const p2b_fri_params& params) {
  std::vector<const F*> cols;
  if (wire_values.size() != circuit.desc().num_wires) throw Error(P2B_ERR_INVALID, "witness column count does not match the circuit");
  for (auto& c : wire_values) {
    if (c.size() != (size_t(1) << circuit.desc().degree_bits)) throw Error(P2B_ERR_INVALID, "witness column length does not match the circuit");
    cols.push_back(c.data());
  }
  std::vector<F> out(p2b_proof_len(circuit.get(), constants_sigmas.get(), &params, public_inputs.size()));
  if (out.empty()) throw Error(P2B_ERR_INVALID, "inconsistent FRI parameters");
  ctx.check(p2b_prove(ctx.get(), circuit.get(), constants_sigmas.get(), circuit_digest.data(), cols.data(), public_inputs.data(),
                      public_inputs.size(), &params, out.data(), out.size()));
  return out;
}

// The asynchronous form (p2b_prove_submit / poll / collect): submit returns once the proof is enqueued and the witness
// has been read, so ONE host thread can keep several contexts busy, or generate the next witness meanwhile (the
// reference does witness generation, then a blocking prove: city_rollup_circuit/src/worker/traits.rs:143-160).
inline size_t prove_submit(const Context& ctx, const CircuitData& circuit, const PolynomialBatch& constants_sigmas,
                           const HashOut& circuit_digest, const std::vector<const F*>& wire_cols,
                           const std::vector<F>& public_inputs, const p2b_fri_params& params) {
  const size_t len = p2b_proof_len(circuit.get(), constants_sigmas.get(), &params, public_inputs.size());
  if (!len) throw Error(P2B_ERR_INVALID, "inconsistent FRI parameters");
  if (wire_cols.size() != circuit.desc().num_wires) throw Error(P2B_ERR_INVALID, "witness column count does not match the circuit");
  ctx.check(p2b_prove_submit(ctx.get(), circuit.get(), constants_sigmas.get(), circuit_digest.data(), wire_cols.data(),
                             public_inputs.data(), public_inputs.size(), &params));
  return len;
}
// p2b_prove_submit_nowait: no wait for the upload — the form for a thread that drives many contexts; pinned witness columns
// stay untouched until prove_upload_poll(ctx) (or the proof has been collected)
inline size_t prove_submit_nowait(const Context& ctx, const CircuitData& circuit, const PolynomialBatch& constants_sigmas,
                                  const HashOut& circuit_digest, const std::vector<const F*>& wire_cols,
                                  const std::vector<F>& public_inputs, const p2b_fri_params& params) {
  const size_t len = p2b_proof_len(circuit.get(), constants_sigmas.get(), &params, public_inputs.size());
  if (!len) throw Error(P2B_ERR_INVALID, "inconsistent FRI parameters");
  if (wire_cols.size() != circuit.desc().num_wires) throw Error(P2B_ERR_INVALID, "witness column count does not match the circuit");
  ctx.check(p2b_prove_submit_nowait(ctx.get(), circuit.get(), constants_sigmas.get(), circuit_digest.data(), wire_cols.data(),
                                    public_inputs.data(), public_inputs.size(), &params));
  return len;
}
inline bool prove_upload_poll(const Context& ctx) {
  const int rc = p2b_prove_upload_poll(ctx.get());
  if (rc < 0) ctx.check(rc);
  return rc == 1;
}
inline bool prove_poll(const Context& ctx) {
  const int rc = p2b_prove_poll(ctx.get());
  if (rc < 0) ctx.check(rc);
  return rc == 1;
}
inline std::vector<F> prove_collect(const Context& ctx, size_t len) {
  std::vector<F> out(len);
  ctx.check(p2b_prove_collect(ctx.get(), out.data(), out.size()));
  return out;
}

// bincode::serialize(&ProofWithPublicInputs) / bincode::deserialize — the bytes the reference's proof store holds
// (city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46,65-72) from / to the flat proof words
inline std::vector<uint8_t> proof_to_bincode(const p2b_proof_shape& shape, const p2b_fri_params& params, const std::vector<F>& words) {
  std::vector<uint8_t> out(p2b_proof_bincode_len(&shape, &params));
  size_t written = 0;
  const int rc = out.empty() ? (int)P2B_ERR_INVALID
                             : p2b_proof_to_bincode(&shape, &params, words.data(), words.size(), out.data(), out.size(), &written);
  if (rc != P2B_OK) throw Error(rc, "proof words do not match the shape");
  out.resize(written);
  return out;
}
inline std::vector<F> proof_from_bincode(const p2b_proof_shape& shape, const p2b_fri_params& params, const std::vector<uint8_t>& bytes) {
  std::vector<F> words(p2b_proof_words(&shape, &params));
  size_t n = 0;
  const int rc = words.empty() ? (int)P2B_ERR_INVALID
                               : p2b_proof_from_bincode(&shape, &params, bytes.data(), bytes.size(), words.data(), words.size(), &n);
  if (rc != P2B_OK) throw Error(rc, "blob is not a proof of this shape");
  return words;
}

}  // namespace plonky2_b200
