//! Raw bindings of include/p2b.h plus the minimal safe wrappers the patched `plonky2` crate uses.
//! Every `extern "C"` item mirrors one declaration of the header; see INTEGRATION.md for the call
//! sites inside plonky2 (fri/oracle.rs, hash/merkle_tree.rs, fri/prover.rs, iop/challenger.rs).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct p2b_ctx { _p: [u8; 0] }
#[repr(C)] pub struct p2b_batch { _p: [u8; 0] }
#[repr(C)] pub struct p2b_tree { _p: [u8; 0] }
#[repr(C)] pub struct p2b_challenger { _p: [u8; 0] }
#[repr(C)] pub struct p2b_circuit { _p: [u8; 0] }

pub const P2B_KEEP_VALUES: u32 = 1;

/// `p2b_gate`: one entry of `common_data.gates` with its selector layout (`selectors_info`).
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct p2b_gate { pub kind: u32, pub p0: u32, pub p1: u32, pub selector_index: u32, pub group_start: u32, pub group_end: u32, pub row: u32 }
/// `p2b_circuit_desc`: the slice of `CommonCircuitData` the prover stages read.
#[repr(C)]
pub struct p2b_circuit_desc {
    pub degree_bits: u32, pub num_wires: u32, pub num_routed_wires: u32, pub num_constants: u32, pub num_selectors: u32,
    pub num_challenges: u32, pub quotient_degree_factor: u32, pub num_partial_products: u32, pub num_gate_constraints: u32,
    pub n_gates: u32, pub gates: *const p2b_gate, pub k_is: *const u64,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct p2b_fri_range { pub oracle: u32, pub first: u32, pub count: u32 }
#[repr(C)]
pub struct p2b_fri_batch { pub point: [u64; 2], pub n_ranges: u32, pub ranges: [p2b_fri_range; 8] }
#[repr(C)]
pub struct p2b_fri_params {
    pub rate_bits: u32, pub cap_height: u32, pub proof_of_work_bits: u32, pub num_query_rounds: u32,
    pub n_layers: u32, pub reduction_arity_bits: [u32; 16],
}

extern "C" {
    pub fn p2b_version() -> c_int;
    pub fn p2b_init(device: c_int, out: *mut *mut p2b_ctx) -> c_int;
    pub fn p2b_set_blocking_sync(ctx: *mut p2b_ctx, on: c_int) -> c_int;
    pub fn p2b_init_on_stream(device: c_int, stream: *mut c_void, out: *mut *mut p2b_ctx) -> c_int;
    pub fn p2b_destroy(ctx: *mut p2b_ctx);
    pub fn p2b_last_error(ctx: *const p2b_ctx) -> *const c_char;
    pub fn p2b_synchronize(ctx: *mut p2b_ctx) -> c_int;
    pub fn p2b_host_alloc(ctx: *mut p2b_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn p2b_host_free(ctx: *mut p2b_ctx, p: *mut c_void) -> c_int;
    pub fn p2b_launch_count(ctx: *const p2b_ctx) -> u64;
    pub fn p2b_timer_start(ctx: *mut p2b_ctx) -> c_int;
    pub fn p2b_timer_stop_ms(ctx: *mut p2b_ctx, ms: *mut f32) -> c_int;
    pub fn p2b_profile_enable(ctx: *mut p2b_ctx, on: c_int) -> c_int;
    pub fn p2b_profile_read(ctx: *mut p2b_ctx, ms: *mut f32, count: *mut u64) -> c_int;

    pub fn p2b_batch_from_values(ctx: *mut p2b_ctx, cols: *const *const u64, n_cols: usize, log_n: u32,
        rate_bits: u32, cap_height: u32, flags: u32, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_from_coeffs(ctx: *mut p2b_ctx, cols: *const *const u64, n_cols: usize, log_n: u32,
        rate_bits: u32, cap_height: u32, flags: u32, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_from_values_dev(ctx: *mut p2b_ctx, d_cols: *const u64, n_cols: usize, log_n: u32,
        rate_bits: u32, cap_height: u32, flags: u32, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_from_coeffs_dev(ctx: *mut p2b_ctx, d_cols: *const u64, n_cols: usize, log_n: u32,
        rate_bits: u32, cap_height: u32, flags: u32, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_free(b: *mut p2b_batch);
    pub fn p2b_batch_n_cols(b: *const p2b_batch) -> usize;
    pub fn p2b_batch_degree_log(b: *const p2b_batch) -> u32;
    pub fn p2b_batch_rate_bits(b: *const p2b_batch) -> u32;
    pub fn p2b_batch_tree(b: *mut p2b_batch) -> *mut p2b_tree;
    pub fn p2b_batch_cap(b: *mut p2b_batch, out: *mut u64) -> c_int;
    pub fn p2b_batch_coeffs(b: *mut p2b_batch, col: usize, out: *mut u64) -> c_int;
    pub fn p2b_batch_leaf(b: *mut p2b_batch, leaf_index: usize, out: *mut u64) -> c_int;
    pub fn p2b_batch_lde_values(b: *mut p2b_batch, index: usize, step: usize, out: *mut u64) -> c_int;
    pub fn p2b_batch_leaves(b: *mut p2b_batch, out: *mut u64) -> c_int;
    pub fn p2b_batch_dev_lde(b: *const p2b_batch) -> *const u64;
    pub fn p2b_batch_dev_coeffs(b: *const p2b_batch) -> *const u64;

    pub fn p2b_merkle_new(ctx: *mut p2b_ctx, leaves: *const u64, n_leaves: usize, leaf_len: usize,
        cap_height: u32, out: *mut *mut p2b_tree) -> c_int;
    pub fn p2b_tree_free(t: *mut p2b_tree);
    pub fn p2b_tree_n_leaves(t: *const p2b_tree) -> usize;
    pub fn p2b_tree_cap_height(t: *const p2b_tree) -> u32;
    pub fn p2b_tree_cap(t: *mut p2b_tree, out: *mut u64) -> c_int;
    pub fn p2b_tree_prove(t: *mut p2b_tree, leaf_index: usize, out: *mut u64) -> c_int;
    pub fn p2b_tree_digests(t: *mut p2b_tree, out: *mut u64) -> c_int;
    pub fn p2b_tree_leaf(t: *mut p2b_tree, leaf_index: usize, out: *mut u64) -> c_int;

    pub fn p2b_poseidon_permute(ctx: *mut p2b_ctx, states: *mut u64, n: usize) -> c_int;
    pub fn p2b_hash_no_pad(ctx: *mut p2b_ctx, input: *const u64, len: usize, out: *mut u64) -> c_int;
    pub fn p2b_two_to_one(ctx: *mut p2b_ctx, left: *const u64, right: *const u64, n: usize, out: *mut u64) -> c_int;

    pub fn p2b_challenger_new(ctx: *mut p2b_ctx, out: *mut *mut p2b_challenger) -> c_int;
    pub fn p2b_challenger_free(c: *mut p2b_challenger);
    pub fn p2b_challenger_observe(c: *mut p2b_challenger, elems: *const u64, n: usize) -> c_int;
    pub fn p2b_challenger_observe_cap(c: *mut p2b_challenger, t: *mut p2b_tree) -> c_int;
    pub fn p2b_challenger_get(c: *mut p2b_challenger, n: usize, out: *mut u64) -> c_int;
    pub fn p2b_challenger_export(c: *mut p2b_challenger, out30: *mut u64) -> c_int;
    pub fn p2b_challenger_import(c: *mut p2b_challenger, in30: *const u64) -> c_int;

    pub fn p2b_fri_commit(ctx: *mut p2b_ctx, coeffs_ext: *const u64, values_ext: *const u64, len: usize,
        arity_bits: *const u32, n_layers: usize, rate_bits: u32, cap_height: u32,
        challenger: *mut p2b_challenger, layers_out: *mut *mut p2b_tree, final_poly_out: *mut u64) -> c_int;
    pub fn p2b_fri_pow(ctx: *mut p2b_ctx, challenger: *mut p2b_challenger, pow_bits: u32, witness_out: *mut u64) -> c_int;

    pub fn p2b_batch_values(b: *mut p2b_batch, col: usize, out: *mut u64) -> c_int;
    pub fn p2b_circuit_new(ctx: *mut p2b_ctx, desc: *const p2b_circuit_desc, out: *mut *mut p2b_circuit) -> c_int;
    pub fn p2b_circuit_free(c: *mut p2b_circuit);
    pub fn p2b_zs_partial_products_commit(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        wires: *const p2b_batch, betas: *const u64, gammas: *const u64, rate_bits: u32, cap_height: u32,
        out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_quotient_commit(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        wires: *const p2b_batch, zs_partial_products: *const p2b_batch, pi_hash: *const u64, betas: *const u64,
        gammas: *const u64, alphas: *const u64, rate_bits: u32, cap_height: u32, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_eval_ext(b: *mut p2b_batch, point: *const u64, first: usize, count: usize, out: *mut u64) -> c_int;
    pub fn p2b_fri_proof_len(oracles: *const *const p2b_batch, n_oracles: usize, params: *const p2b_fri_params) -> usize;
    pub fn p2b_prove_openings(ctx: *mut p2b_ctx, oracles: *const *const p2b_batch, n_oracles: usize,
        batches: *const p2b_fri_batch, n_batches: usize, challenger: *mut p2b_challenger, params: *const p2b_fri_params,
        proof_out: *mut u64, proof_cap: usize) -> c_int;
    pub fn p2b_proof_len(circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch, params: *const p2b_fri_params,
        n_public_inputs: usize) -> usize;
    pub fn p2b_prove(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        circuit_digest: *const u64, wire_cols: *const *const u64, public_inputs: *const u64, n_public_inputs: usize,
        params: *const p2b_fri_params, proof_out: *mut u64, proof_cap: usize) -> c_int;

    pub fn p2b_proof_words(shape: *const p2b_proof_shape, params: *const p2b_fri_params) -> usize;
    pub fn p2b_proof_bincode_len(shape: *const p2b_proof_shape, params: *const p2b_fri_params) -> usize;
    pub fn p2b_proof_to_bincode(shape: *const p2b_proof_shape, params: *const p2b_fri_params, words: *const u64,
        n_words: usize, out: *mut u8, out_cap: usize, written: *mut usize) -> c_int;
    pub fn p2b_proof_from_bincode(shape: *const p2b_proof_shape, params: *const p2b_fri_params, bytes: *const u8,
        n_bytes: usize, words_out: *mut u64, words_cap: usize, n_words: *mut usize) -> c_int;
}

/// `p2b_proof_shape`: what of `CommonCircuitData` fixes the layout of a proof
/// (city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct p2b_proof_shape {
    pub degree_bits: u32,
    pub num_constants: u32, pub num_routed_wires: u32, pub num_wires: u32,
    pub num_challenges: u32, pub num_partial_products: u32, pub quotient_degree_factor: u32,
    pub constants_sigmas_cap_height: u32,
    pub n_public_inputs: u32,
}

/// The bytes `bincode::serialize(&proof_with_pis)` produces (what the proof store keeps,
/// city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46) straight from the proof words.
pub fn proof_to_bincode(shape: &p2b_proof_shape, params: &p2b_fri_params, words: &[u64]) -> Result<Vec<u8>, P2bError> {
    let cap = unsafe { p2b_proof_bincode_len(shape, params) };
    let mut out = vec![0u8; cap];
    let mut written = 0usize;
    let rc = unsafe { p2b_proof_to_bincode(shape, params, words.as_ptr(), words.len(), out.as_mut_ptr(), cap, &mut written) };
    if cap == 0 || rc != 0 { return Err(P2bError { code: rc, message: "proof words do not match the shape".into() }); }
    out.truncate(written);
    Ok(out)
}

/// Error type the patched plonky2 converts into `anyhow::Error` (the reference propagates it with `?`
/// up to `SimpleActorWorker::process_job`, city_rollup_core_worker/src/actors/simple.rs:83).
#[derive(Debug)]
pub struct P2bError { pub code: i32, pub message: String }
impl std::fmt::Display for P2bError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "p2b error {}: {}", self.code, self.message) }
}
impl std::error::Error for P2bError {}

/// One context per worker thread (a p2b_ctx is not thread-safe).
pub struct Context(pub *mut p2b_ctx);
unsafe impl Send for Context {}
impl Context {
    pub fn new(device: i32) -> Result<Self, P2bError> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { p2b_init(device, &mut h) };
        if rc != 0 { return Err(P2bError { code: rc, message: last_error(std::ptr::null()) }); }
        Ok(Context(h))
    }
    pub fn check(&self, rc: c_int) -> Result<(), P2bError> {
        if rc == 0 { Ok(()) } else { Err(P2bError { code: rc, message: last_error(self.0) }) }
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { p2b_destroy(self.0) } } }

pub fn last_error(ctx: *const p2b_ctx) -> String {
    unsafe { CStr::from_ptr(p2b_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Owning handle of a device-resident PolynomialBatch.
pub struct Batch(pub *mut p2b_batch);
unsafe impl Send for Batch {}
impl Drop for Batch { fn drop(&mut self) { unsafe { p2b_batch_free(self.0) } } }

/// `PolynomialBatch::from_values` for `F = GoldilocksField` (a transparent u64): one pointer per column.
pub fn batch_from_values(ctx: &Context, cols: &[&[u64]], rate_bits: usize, cap_height: usize) -> Result<Batch, P2bError> {
    let n = cols[0].len();
    assert!(n.is_power_of_two() && cols.iter().all(|c| c.len() == n));
    let ptrs: Vec<*const u64> = cols.iter().map(|c| c.as_ptr()).collect();
    let mut h = std::ptr::null_mut();
    ctx.check(unsafe {
        p2b_batch_from_values(ctx.0, ptrs.as_ptr(), ptrs.len(), n.trailing_zeros(), rate_bits as u32, cap_height as u32, 0, &mut h)
    })?;
    Ok(Batch(h))
}


/// Owning handle of the uploaded circuit description (built once per `CircuitData`, next to
/// `prover_only.constants_sigmas_commitment`).
pub struct Circuit(pub *mut p2b_circuit);
unsafe impl Send for Circuit {}
impl Drop for Circuit { fn drop(&mut self) { unsafe { p2b_circuit_free(self.0) } } }

/// `prove_with_partition_witness` after witness generation: the flat proof words in `ProofWithPublicInputs`
/// field order (see include/p2b.h); the patched plonky2 re-wraps them into `ProofWithPublicInputs<F, C, 2>`.
pub fn prove(ctx: &Context, circuit: &Circuit, constants_sigmas: &Batch, circuit_digest: &[u64; 4],
             wire_values: &[&[u64]], public_inputs: &[u64], params: &p2b_fri_params) -> Result<Vec<u64>, P2bError> {
    let ptrs: Vec<*const u64> = wire_values.iter().map(|c| c.as_ptr()).collect();
    let len = unsafe { p2b_proof_len(circuit.0, constants_sigmas.0, params, public_inputs.len()) };
    let mut out = vec![0u64; len];
    ctx.check(unsafe {
        p2b_prove(ctx.0, circuit.0, constants_sigmas.0, circuit_digest.as_ptr(), ptrs.as_ptr(), public_inputs.as_ptr(),
                  public_inputs.len(), params, out.as_mut_ptr(), len)
    })?;
    Ok(out)
}
