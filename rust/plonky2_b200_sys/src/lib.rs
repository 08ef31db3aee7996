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
    pub fn p2b_set_latency_mode(ctx: *mut p2b_ctx, on: c_int) -> c_int;
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

    pub fn p2b_prove_dev(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        circuit_digest: *const u64, d_wire_values: *const u64, public_inputs: *const u64, n_public_inputs: usize,
        params: *const p2b_fri_params, proof_out: *mut u64, proof_cap: usize) -> c_int;
    pub fn p2b_prove_submit(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        circuit_digest: *const u64, wire_cols: *const *const u64, public_inputs: *const u64, n_public_inputs: usize,
        params: *const p2b_fri_params) -> c_int;
    pub fn p2b_prove_poll(ctx: *mut p2b_ctx) -> c_int;
    pub fn p2b_prove_submit_nowait(ctx: *mut p2b_ctx, circuit: *const p2b_circuit, constants_sigmas: *const p2b_batch,
        circuit_digest: *const u64, wire_cols: *const *const u64, public_inputs: *const u64, n_public_inputs: usize,
        params: *const p2b_fri_params) -> c_int;
    pub fn p2b_prove_upload_poll(ctx: *mut p2b_ctx) -> c_int;
    pub fn p2b_prove_collect(ctx: *mut p2b_ctx, proof_out: *mut u64, proof_cap: usize) -> c_int;
    pub fn p2b_plan_info(ctx: *mut p2b_ctx, n_ready: *mut u32, n_seen: *mut u32, n_failed: *mut u32, kernels_per_launch: *mut u32) -> c_int;
    pub fn p2b_batch_attach(ctx: *mut p2b_ctx, src: *const p2b_batch, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_export_len(b: *const p2b_batch) -> usize;
    pub fn p2b_batch_export(b: *mut p2b_batch, out: *mut u8, out_cap: usize, written: *mut usize) -> c_int;
    pub fn p2b_batch_import(ctx: *mut p2b_ctx, bytes: *const u8, n_bytes: usize, out: *mut *mut p2b_batch) -> c_int;
    pub fn p2b_batch_lde_col(b: *mut p2b_batch, col: usize, out: *mut u64) -> c_int;
    pub fn p2b_timer_span_ms(first: *mut p2b_ctx, last: *mut p2b_ctx, ms: *mut f32) -> c_int;

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

/// One context per worker thread (a p2b_ctx is not thread-safe).  `Batch` and `Circuit` borrow their context
/// (`&'ctx Context`), so the borrow checker enforces the header's rule that no handle outlives its context.
pub struct Context(*mut p2b_ctx);
unsafe impl Send for Context {}
impl Context {
    pub fn new(device: i32) -> Result<Self, P2bError> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { p2b_init(device, &mut h) };
        if rc != 0 { return Err(P2bError { code: rc, message: last_error(std::ptr::null()) }); }
        Ok(Context(h))
    }
    pub fn raw(&self) -> *mut p2b_ctx { self.0 }
    /// Waiting threads sleep instead of spinning (several proving threads per host core).
    pub fn set_blocking_sync(&self, on: bool) -> Result<(), P2bError> { self.check(unsafe { p2b_set_blocking_sync(self.0, on as c_int) }) }
    /// This worker has the GPU to itself (one job at a time, `simple.rs:32-56` as the only process on the device):
    /// shorter launch chains for more work.  The default is throughput mode.
    pub fn set_latency_mode(&self, on: bool) -> Result<(), P2bError> { self.check(unsafe { p2b_set_latency_mode(self.0, on as c_int) }) }
    pub fn check(&self, rc: c_int) -> Result<(), P2bError> {
        if rc == 0 { Ok(()) } else { Err(P2bError { code: rc, message: last_error(self.0) }) }
    }
    fn invalid(msg: impl Into<String>) -> P2bError { P2bError { code: -1, message: msg.into() } }
}
impl Drop for Context { fn drop(&mut self) { unsafe { p2b_destroy(self.0) } } }

pub fn last_error(ctx: *const p2b_ctx) -> String {
    unsafe { CStr::from_ptr(p2b_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Columns of one power-of-two length: the checks every column-taking wrapper needs before it hands raw pointers to C.
fn column_ptrs(cols: &[&[u64]], expect_cols: Option<usize>, expect_len: Option<usize>) -> Result<(Vec<*const u64>, u32), P2bError> {
    if cols.is_empty() { return Err(Context::invalid("no columns")); }
    if let Some(k) = expect_cols {
        if cols.len() != k { return Err(Context::invalid(format!("{} columns given, the circuit has {}", cols.len(), k))); }
    }
    let n = cols[0].len();
    if n == 0 || !n.is_power_of_two() { return Err(Context::invalid("column length must be a power of two")); }
    if let Some(l) = expect_len {
        if n != l { return Err(Context::invalid(format!("columns of {} values given, the circuit has 2^{} rows", n, l.trailing_zeros()))); }
    }
    if cols.iter().any(|c| c.len() != n) { return Err(Context::invalid("columns of different lengths")); }
    Ok((cols.iter().map(|c| c.as_ptr()).collect(), n.trailing_zeros()))
}

/// Owning handle of a device-resident PolynomialBatch.
pub struct Batch<'ctx> { h: *mut p2b_batch, ctx: &'ctx Context, n_cols: usize, log_n: u32 }
unsafe impl<'ctx> Send for Batch<'ctx> {}
impl<'ctx> Drop for Batch<'ctx> { fn drop(&mut self) { unsafe { p2b_batch_free(self.h) } } }
impl<'ctx> Batch<'ctx> {
    pub fn raw(&self) -> *mut p2b_batch { self.h }
    pub fn n_cols(&self) -> usize { self.n_cols }
    pub fn degree_log(&self) -> u32 { self.log_n }
    fn wrap(ctx: &'ctx Context, h: *mut p2b_batch) -> Self {
        let (n_cols, log_n) = unsafe { (p2b_batch_n_cols(h), p2b_batch_degree_log(h)) };
        Batch { h, ctx, n_cols, log_n }
    }
    /// `PolynomialBatch::from_values` for `F = GoldilocksField` (a transparent u64): one pointer per column.
    /// `keep_values` keeps the values on H in HBM (needed for `prover_data.constants_sigmas_commitment` and the wires).
    pub fn from_values(ctx: &'ctx Context, cols: &[&[u64]], rate_bits: usize, cap_height: usize, keep_values: bool) -> Result<Self, P2bError> {
        let (ptrs, log_n) = column_ptrs(cols, None, None)?;
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe {
            p2b_batch_from_values(ctx.0, ptrs.as_ptr(), ptrs.len(), log_n, rate_bits as u32, cap_height as u32,
                                  if keep_values { P2B_KEEP_VALUES } else { 0 }, &mut h)
        })?;
        Ok(Self::wrap(ctx, h))
    }
    /// A read-only view of this batch for another context of the same device: the worker threads of a GPU share one
    /// device copy of a circuit's constants|sigmas batch.  The view cannot outlive `self` ('src: 'ctx2 is enforced).
    pub fn attach<'c2>(&'c2 self, ctx: &'c2 Context) -> Result<Batch<'c2>, P2bError> {
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { p2b_batch_attach(ctx.0, self.h, &mut h) })?;
        Ok(Batch::wrap(ctx, h))
    }
    /// Bytes a later process re-imports instead of re-running `CircuitBuilder::build`'s commitment.
    pub fn export(&self) -> Result<Vec<u8>, P2bError> {
        let cap = unsafe { p2b_batch_export_len(self.h) };
        let mut out = vec![0u8; cap];
        let mut written = 0usize;
        self.ctx.check(unsafe { p2b_batch_export(self.h, out.as_mut_ptr(), cap, &mut written) })?;
        out.truncate(written);
        Ok(out)
    }
    pub fn import(ctx: &'ctx Context, bytes: &[u8]) -> Result<Self, P2bError> {
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { p2b_batch_import(ctx.0, bytes.as_ptr(), bytes.len(), &mut h) })?;
        Ok(Self::wrap(ctx, h))
    }
}

/// Owning handle of the uploaded circuit description (built once per `CircuitData`, next to
/// `prover_only.constants_sigmas_commitment`).  Remembers the witness shape so that `prove` can validate it.
pub struct Circuit<'ctx> { h: *mut p2b_circuit, ctx: &'ctx Context, num_wires: usize, degree_bits: u32 }
unsafe impl<'ctx> Send for Circuit<'ctx> {}
impl<'ctx> Drop for Circuit<'ctx> { fn drop(&mut self) { unsafe { p2b_circuit_free(self.h) } } }
impl<'ctx> Circuit<'ctx> {
    /// `gates` / `k_is` are the slices `desc` would point to; the pointers inside the C struct are filled in here.
    pub fn new(ctx: &'ctx Context, mut desc: p2b_circuit_desc, gates: &[p2b_gate], k_is: &[u64]) -> Result<Self, P2bError> {
        if k_is.len() != desc.num_routed_wires as usize { return Err(Context::invalid("k_is must have num_routed_wires entries")); }
        desc.n_gates = gates.len() as u32;
        desc.gates = gates.as_ptr();
        desc.k_is = k_is.as_ptr();
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { p2b_circuit_new(ctx.0, &desc, &mut h) })?;
        Ok(Circuit { h, ctx, num_wires: desc.num_wires as usize, degree_bits: desc.degree_bits })
    }
    pub fn raw(&self) -> *mut p2b_circuit { self.h }
}

fn same_ctx(ctx: &Context, circuit: &Circuit, cs: &Batch) -> Result<(), P2bError> {
    if !std::ptr::eq(circuit.ctx, ctx) || !std::ptr::eq(cs.ctx, ctx) { return Err(Context::invalid("handle of another context")); }
    Ok(())
}

/// `prove_with_partition_witness` after witness generation: the flat proof words in `ProofWithPublicInputs`
/// field order (see include/p2b.h); the patched plonky2 re-wraps them into `ProofWithPublicInputs<F, C, 2>`.
/// The witness shape is checked against the circuit here: the C side reads `num_wires` pointers of 2^degree_bits words.
pub fn prove(ctx: &Context, circuit: &Circuit, constants_sigmas: &Batch, circuit_digest: &[u64; 4],
             wire_values: &[&[u64]], public_inputs: &[u64], params: &p2b_fri_params) -> Result<Vec<u64>, P2bError> {
    same_ctx(ctx, circuit, constants_sigmas)?;
    let (ptrs, _) = column_ptrs(wire_values, Some(circuit.num_wires), Some(1usize << circuit.degree_bits))?;
    let len = unsafe { p2b_proof_len(circuit.h, constants_sigmas.h, params, public_inputs.len()) };
    if len == 0 { return Err(Context::invalid("inconsistent FRI parameters")); }
    let mut out = vec![0u64; len];
    ctx.check(unsafe {
        p2b_prove(ctx.0, circuit.h, constants_sigmas.h, circuit_digest.as_ptr(), ptrs.as_ptr(), public_inputs.as_ptr(),
                  public_inputs.len(), params, out.as_mut_ptr(), len)
    })?;
    Ok(out)
}

/// A proof in flight on a context (`p2b_prove_submit`): the witness buffers are free again as soon as `submit`
/// returns, so the worker thread generates the next witness while the GPU proves this one
/// (city_rollup_circuit/src/worker/traits.rs:143-160 is witness generation followed by a blocking prove).
/// One proof may be pending per context: a second `prove_submit` before `collect` fails with P2B_ERR_INVALID.
pub struct PendingProof<'a> { ctx: &'a Context, len: usize }
pub fn prove_submit<'a>(ctx: &'a Context, circuit: &Circuit, constants_sigmas: &Batch, circuit_digest: &[u64; 4],
                        wire_values: &[&[u64]], public_inputs: &[u64], params: &p2b_fri_params) -> Result<PendingProof<'a>, P2bError> {
    same_ctx(ctx, circuit, constants_sigmas)?;
    let (ptrs, _) = column_ptrs(wire_values, Some(circuit.num_wires), Some(1usize << circuit.degree_bits))?;
    let len = unsafe { p2b_proof_len(circuit.h, constants_sigmas.h, params, public_inputs.len()) };
    if len == 0 { return Err(Context::invalid("inconsistent FRI parameters")); }
    let rc = unsafe {
        p2b_prove_submit(ctx.0, circuit.h, constants_sigmas.h, circuit_digest.as_ptr(), ptrs.as_ptr(), public_inputs.as_ptr(),
                         public_inputs.len(), params)
    };
    ctx.check(rc)?;
    Ok(PendingProof { ctx, len })
}
/// `p2b_prove_submit_nowait`: no wait for the upload of pinned witness columns — the form for ONE thread that keeps many
/// contexts busy (an upload can wait milliseconds behind other contexts' kernels).  The returned proof borrows the
/// witness until it is collected, so safe code cannot refill the columns while the DMA may still read them.
pub struct PendingProofBorrowing<'a, 'w> { inner: PendingProof<'a>, _witness: core::marker::PhantomData<&'w [u64]> }
pub fn prove_submit_nowait<'a, 'w>(ctx: &'a Context, circuit: &Circuit, constants_sigmas: &Batch, circuit_digest: &[u64; 4],
                                   wire_values: &'w [&'w [u64]], public_inputs: &[u64], params: &p2b_fri_params)
                                   -> Result<PendingProofBorrowing<'a, 'w>, P2bError> {
    same_ctx(ctx, circuit, constants_sigmas)?;
    let (ptrs, _) = column_ptrs(wire_values, Some(circuit.num_wires), Some(1usize << circuit.degree_bits))?;
    let len = unsafe { p2b_proof_len(circuit.h, constants_sigmas.h, params, public_inputs.len()) };
    if len == 0 { return Err(Context::invalid("inconsistent FRI parameters")); }
    let rc = unsafe {
        p2b_prove_submit_nowait(ctx.0, circuit.h, constants_sigmas.h, circuit_digest.as_ptr(), ptrs.as_ptr(),
                                public_inputs.as_ptr(), public_inputs.len(), params)
    };
    ctx.check(rc)?;
    Ok(PendingProofBorrowing { inner: PendingProof { ctx, len }, _witness: core::marker::PhantomData })
}
impl<'a, 'w> PendingProofBorrowing<'a, 'w> {
    /// true once the witness has been read by the device
    pub fn upload_finished(&self) -> Result<bool, P2bError> {
        let rc = unsafe { p2b_prove_upload_poll(self.inner.ctx.0) };
        if rc < 0 { self.inner.ctx.check(rc)?; }
        Ok(rc == 1)
    }
    pub fn is_finished(&self) -> Result<bool, P2bError> { self.inner.is_finished() }
    pub fn collect(self) -> Result<Vec<u64>, P2bError> { self.inner.collect() }
}
impl<'a> PendingProof<'a> {
    pub fn is_finished(&self) -> Result<bool, P2bError> {
        let rc = unsafe { p2b_prove_poll(self.ctx.0) };
        if rc < 0 { self.ctx.check(rc)?; }
        Ok(rc == 1)
    }
    pub fn collect(self) -> Result<Vec<u64>, P2bError> {
        let mut out = vec![0u64; self.len];
        let rc = unsafe { p2b_prove_collect(self.ctx.0, out.as_mut_ptr(), self.len) };
        self.ctx.check(rc)?;
        Ok(out)
    }
}
