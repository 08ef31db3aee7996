"""Host-side mirror of the plonky2 prover surface City Rollup's workers call, backed by the CUDA
library (libp2b.so) through its C ABI.

The reference is Rust and no Rust toolchain exists in this image, so this Python layer plays the part of
the patched `plonky2` crate (INTEGRATION.md shows the Rust binding): same names, argument meaning and
error behaviour as plonky2 0.2.2 —

    PolynomialBatch.from_values / from_coeffs / get_lde_values   (fri/oracle.rs)
    MerkleTree.new / prove / cap / digests / get                 (hash/merkle_tree.rs)
    Challenger.observe_* / get_*                                  (iop/challenger.rs)
    fri_committed_trees / fri_proof_of_work                      (fri/prover.rs)
    all_wires_permutation_partial_products / compute_quotient_polys  (plonk/prover.rs), on a
    CommonCircuitData-shaped circuit description (CircuitData below)

reached in the reference only via `circuit_data.prove(pw)` (e.g.
city_common_circuit/src/proof_minifier/pm_core.rs:151).  Field elements are numpy uint64.
Errors surface as P2BError (the analogue of the `anyhow::Error` the worker loop propagates,
city_rollup_core_worker/src/actors/simple.rs:83); nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib

u64p = _lib.u64p


class P2BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"p2b error {code}: {msg}")
        self.code = code


def _ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "expected a contiguous uint64 array"
    return a.ctypes.data_as(u64p)


class Context:
    """One CUDA device + stream (p2b_ctx).  Not thread-safe; create one per worker thread."""

    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.p2b_init(device, C.byref(h))
        else:
            rc = self.lib.p2b_init_on_stream(device, C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise P2BError(rc, self.lib.p2b_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise P2BError(rc, self.lib.p2b_last_error(self.h).decode())

    def set_blocking_sync(self, on=True):
        """sleep instead of spinning while waiting for the device (many proving threads per host core)"""
        self.check(self.lib.p2b_set_blocking_sync(self.h, 1 if on else 0))

    def set_latency_mode(self, on=True):
        """p2b_set_latency_mode: this worker has the GPU to itself (one proof at a time) — shorter launch chains for more
        work; the default is throughput mode (many proofs in flight per GPU).  Same results."""
        self.check(self.lib.p2b_set_latency_mode(self.h, 1 if on else 0))

    def synchronize(self):
        self.check(self.lib.p2b_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.p2b_launch_count(self.h))

    def timer_start(self):
        self.check(self.lib.p2b_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        self.check(self.lib.p2b_timer_stop_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def timer_span_ms(self, last):
        """ms from this context's timer_start to `last`'s timer_stop_ms (contexts of one device working side by side)"""
        ms = C.c_float()
        last.check(self.lib.p2b_timer_span_ms(self.h, last.h, C.byref(ms)))
        return float(ms.value)

    def plan_info(self):
        """prove plans (captured CUDA graphs) of this context: dict(ready, seen, failed, kernels_per_launch, note)"""
        v = [C.c_uint32() for _ in range(4)]
        self.check(self.lib.p2b_plan_info(self.h, *[C.byref(x) for x in v]))
        note = self.lib.p2b_last_error(self.h).decode() if v[2].value else ""
        return dict(ready=v[0].value, seen=v[1].value, failed=v[2].value, kernels_per_launch=v[3].value, note=note)

    STAGES = ("h2d", "intt", "lde", "leaf_hash", "tree_levels", "fri_fold_ntt", "transcript", "other")

    def profile_enable(self, on=True):
        self.check(self.lib.p2b_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """-> ({stage: ms}, {stage: launches}) accumulated since the previous read"""
        ms = (C.c_float * 8)()
        cnt = (C.c_uint64 * 8)()
        self.check(self.lib.p2b_profile_read(self.h, ms, cnt))
        return ({s: float(ms[i]) for i, s in enumerate(self.STAGES)},
                {s: int(cnt[i]) for i, s in enumerate(self.STAGES)})

    def pinned_empty(self, shape):
        """uint64 array backed by pinned host memory (freed with the context)."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        self.check(self.lib.p2b_host_alloc(self.h, n * 8, C.byref(p)))
        buf = (C.c_uint64 * n).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint64).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return a

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.lib.p2b_host_free(self.h, p)
            self._pinned = []
            self.lib.p2b_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Poseidon utilities (PoseidonHash) ----
    def poseidon_permute(self, states):
        s = np.ascontiguousarray(np.array(states, dtype=np.uint64)).reshape(-1, 12).copy()
        self.check(self.lib.p2b_poseidon_permute(self.h, _ptr(s), s.shape[0]))
        return s

    def hash_no_pad(self, x):
        x = np.ascontiguousarray(np.array(x, dtype=np.uint64)).reshape(-1)
        o = np.zeros(4, np.uint64)
        xp = _ptr(x) if x.size else None
        self.check(self.lib.p2b_hash_no_pad(self.h, xp, x.size, _ptr(o)))
        return o

    def two_to_one(self, left, right):
        l = np.ascontiguousarray(np.array(left, dtype=np.uint64)).reshape(-1, 4)
        r = np.ascontiguousarray(np.array(right, dtype=np.uint64)).reshape(-1, 4)
        o = np.zeros_like(l)
        self.check(self.lib.p2b_two_to_one(self.h, _ptr(l), _ptr(r), l.shape[0], _ptr(o)))
        return o


class MerkleTree:
    """plonky2::hash::merkle_tree::MerkleTree<GoldilocksField, PoseidonHash> (device resident)."""

    def __init__(self, ctx, handle, owner=None):
        self.ctx, self.h, self._owner = ctx, handle, owner
        lib = ctx.lib
        self.n_leaves = int(lib.p2b_tree_n_leaves(handle))
        self.cap_height = int(lib.p2b_tree_cap_height(handle))

    @classmethod
    def new(cls, ctx, leaves, cap_height):
        """MerkleTree::new(leaves: Vec<Vec<F>>, cap_height)"""
        leaves = np.ascontiguousarray(np.array(leaves, dtype=np.uint64))
        if leaves.ndim != 2:
            raise ValueError("leaves must be (n_leaves, leaf_len)")
        h = C.c_void_p()
        lp = _ptr(leaves) if leaves.size else _ptr(np.zeros(1, np.uint64))
        ctx.check(ctx.lib.p2b_merkle_new(ctx.h, lp, leaves.shape[0], leaves.shape[1], cap_height, C.byref(h)))
        t = cls(ctx, h)
        t.leaf_len = leaves.shape[1]
        return t

    @property
    def cap(self):
        o = np.zeros((1 << self.cap_height, 4), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_tree_cap(self.h, _ptr(o)))
        return o

    @property
    def digests(self):
        """tree.digests in plonky2's interleaved layout"""
        n = 2 * (self.n_leaves - (1 << self.cap_height))
        o = np.zeros((max(n, 1), 4), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_tree_digests(self.h, _ptr(o)))
        return o[:n]

    def prove(self, leaf_index):
        """MerkleTree::prove(leaf_index).siblings"""
        L = self.n_leaves.bit_length() - 1 - self.cap_height
        o = np.zeros((max(L, 1), 4), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_tree_prove(self.h, leaf_index, _ptr(o)))
        return o[:L]

    def get(self, leaf_index, leaf_len=None):
        """MerkleTree::get(leaf_index)"""
        n = leaf_len if leaf_len is not None else self.leaf_len
        o = np.zeros(max(n, 1), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_tree_leaf(self.h, leaf_index, _ptr(o)))
        return o[:n]

    def free(self):
        if self.h and self._owner is None:
            self.ctx.lib.p2b_tree_free(self.h)
        self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


class PolynomialBatch:
    """plonky2::fri::oracle::PolynomialBatch<GoldilocksField, PoseidonGoldilocksConfig, 2>"""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        lib = ctx.lib
        self.n_cols = int(lib.p2b_batch_n_cols(handle))
        self.degree_log = int(lib.p2b_batch_degree_log(handle))
        self.rate_bits = int(lib.p2b_batch_rate_bits(handle))
        self.merkle_tree = MerkleTree(ctx, C.c_void_p(lib.p2b_batch_tree(handle)), owner=self)
        self.merkle_tree.leaf_len = self.n_cols

    @staticmethod
    def _cols(values):
        """columns -> (objects to keep alive, const uint64_t *const cols[], log2 length, number of columns).  This sits inside every
        from_values / prove call, so the pointer table is built from raw addresses (numpy's `.ctypes` costs ~7 us per
        column: 0.9 ms for a 135-column witness, 10 % of a commit)."""
        if isinstance(values, np.ndarray) and values.ndim == 2 and values.dtype == np.uint64 and values.flags.c_contiguous:
            n_cols, n = values.shape
            if n_cols == 0:
                raise ValueError("empty batch")
            addrs = values.__array_interface__["data"][0] + np.arange(n_cols, dtype=np.uint64) * np.uint64(8 * n)
            cols = [values]
        else:
            cols = [c if (type(c) is np.ndarray and c.dtype == np.uint64 and c.ndim == 1 and c.flags.c_contiguous)
                    else np.ascontiguousarray(np.asarray(c, dtype=np.uint64)).reshape(-1) for c in values]
            if not cols:
                raise ValueError("empty batch")
            n = cols[0].size
            if any(c.size != n for c in cols):
                raise ValueError("all columns must have the same power-of-two length")
            addrs = np.fromiter((c.__array_interface__["data"][0] for c in cols), dtype=np.uint64, count=len(cols))
        if n == 0 or n & (n - 1):
            raise ValueError("all columns must have the same power-of-two length")
        ptrs = C.cast(addrs.ctypes.data, C.POINTER(u64p))
        return (cols, addrs), ptrs, n.bit_length() - 1, int(addrs.size)

    KEEP_VALUES = 1  # P2B_KEEP_VALUES
    BLINDING = 2     # rejected by the library: no worker circuit is zero-knowledge

    @classmethod
    def from_values(cls, ctx, values, rate_bits, blinding, cap_height, timing=None, fft_root_table=None,
                    keep_values=False):
        """PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table);
        keep_values keeps the values on H in HBM (the witness columns / sigmas the Z stage reads again)"""
        keep, ptrs, log_n, n_cols = cls._cols(values)
        h = C.c_void_p()
        flags = (cls.BLINDING if blinding else 0) | (cls.KEEP_VALUES if keep_values else 0)
        ctx.check(ctx.lib.p2b_batch_from_values(ctx.h, ptrs, n_cols, log_n, rate_bits, cap_height, flags,
                                                C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_coeffs(cls, ctx, polynomials, rate_bits, blinding, cap_height, timing=None, fft_root_table=None):
        """PolynomialBatch::from_coeffs(polynomials, rate_bits, blinding, cap_height, timing, fft_root_table)"""
        keep, ptrs, log_n, n_cols = cls._cols(polynomials)
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_batch_from_coeffs(ctx.h, ptrs, n_cols, log_n, rate_bits, cap_height,
                                                cls.BLINDING if blinding else 0, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_values_device(cls, ctx, dev_ptr, n_cols, log_n, rate_bits, cap_height):
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_batch_from_values_dev(ctx.h, C.c_void_p(dev_ptr), n_cols, log_n, rate_bits,
                                                    cap_height, 0, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_coeffs_device(cls, ctx, dev_ptr, n_cols, log_n, rate_bits, cap_height):
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_batch_from_coeffs_dev(ctx.h, C.c_void_p(dev_ptr), n_cols, log_n, rate_bits,
                                                    cap_height, 0, C.byref(h)))
        return cls(ctx, h)

    def attach(self, ctx):
        """a read-only view of this batch for another context of the same device (p2b_batch_attach): the contexts of a
        GPU share one device copy of a circuit's constants|sigmas batch.  This batch must outlive the view."""
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_batch_attach(ctx.h, self.h, C.byref(h)))
        v = PolynomialBatch(ctx, h)
        v._source = self
        return v

    def export(self):
        """serialised form (bytes): header, cap, coefficients (+ kept values) — p2b_batch_export"""
        n = int(self.ctx.lib.p2b_batch_export_len(self.h))
        buf = np.zeros(n, np.uint8)
        w = C.c_size_t()
        self.ctx.check(self.ctx.lib.p2b_batch_export(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(w)))
        return buf[: w.value].tobytes()

    @classmethod
    def import_(cls, ctx, blob):
        """p2b_batch_import: LDE and Merkle tree recomputed on the device, cap checked against the stored one"""
        a = np.frombuffer(blob, dtype=np.uint8)
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_batch_import(ctx.h, a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size, C.byref(h)))
        return cls(ctx, h)

    @property
    def cap(self):
        return self.merkle_tree.cap

    def coeffs(self, col):
        """batch.polynomials[col].coeffs"""
        o = np.zeros(1 << self.degree_log, np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_coeffs(self.h, col, _ptr(o)))
        return o

    def values(self, col):
        """the values on H this batch was built from (keep_values=True)"""
        o = np.zeros(1 << self.degree_log, np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_values(self.h, col, _ptr(o)))
        return o

    def eval_ext(self, point, first=0, count=None):
        """polynomials[first .. first+count] evaluated at the extension point (OpeningSet's eval_commitment)
        -> (count, 2) array"""
        count = self.n_cols - first if count is None else count
        o = np.zeros((max(count, 1), 2), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_eval_ext(self.h, _ptr(_felts(point)), first, count, _ptr(o)))
        return o[:count]

    def lde_col(self, col):
        """the whole LDE of polynomial `col` in leaf order (column `col` of batch.merkle_tree.leaves)"""
        o = np.zeros(1 << (self.degree_log + self.rate_bits), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_lde_col(self.h, col, _ptr(o)))
        return o

    def get_lde_values(self, index, step=1):
        """PolynomialBatch::get_lde_values(index, step)"""
        o = np.zeros(self.n_cols, np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_lde_values(self.h, index, step, _ptr(o)))
        return o

    def leaf(self, leaf_index):
        o = np.zeros(self.n_cols, np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_leaf(self.h, leaf_index, _ptr(o)))
        return o

    def leaves(self):
        """batch.merkle_tree.leaves as an (n_leaves, n_cols) array"""
        N = 1 << (self.degree_log + self.rate_bits)
        o = np.zeros((N, self.n_cols), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_batch_leaves(self.h, _ptr(o)))
        return o

    def free(self):
        if self.h:
            self.merkle_tree.h = None
            self.ctx.lib.p2b_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


class Challenger:
    """plonky2::iop::challenger::Challenger<GoldilocksField, PoseidonHash>, state resident in HBM"""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_challenger_new(ctx.h, C.byref(h)))
        self.h = h

    def observe_elements(self, elems):
        e = np.ascontiguousarray(np.array(elems, dtype=np.uint64)).reshape(-1)
        if e.size:
            self.ctx.check(self.ctx.lib.p2b_challenger_observe(self.h, _ptr(e), e.size))

    observe_element = lambda self, e: self.observe_elements([e])
    observe_hash = observe_elements
    observe_extension_elements = observe_elements

    def observe_cap(self, tree_or_batch):
        t = tree_or_batch.merkle_tree if isinstance(tree_or_batch, PolynomialBatch) else tree_or_batch
        self.ctx.check(self.ctx.lib.p2b_challenger_observe_cap(self.h, t.h))

    def get_n_challenges(self, n):
        o = np.zeros(max(n, 1), np.uint64)
        self.ctx.check(self.ctx.lib.p2b_challenger_get(self.h, n, _ptr(o)))
        return [int(x) for x in o[:n]]

    def get_challenge(self):
        return self.get_n_challenges(1)[0]

    def get_extension_challenge(self):
        return self.get_n_challenges(2)

    def export_state(self):
        o = np.zeros(30, np.uint64)
        self.ctx.check(self.ctx.lib.p2b_challenger_export(self.h, _ptr(o)))
        return o

    def import_state(self, s):
        s = np.ascontiguousarray(np.array(s, dtype=np.uint64))
        self.ctx.check(self.ctx.lib.p2b_challenger_import(self.h, _ptr(s)))

    def free(self):
        if self.h:
            self.ctx.lib.p2b_challenger_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def fri_committed_trees(ctx, coeffs, values, challenger, reduction_arity_bits, rate_bits=3, cap_height=4):
    """fri::prover::fri_committed_trees(coeffs, values, challenger, fri_params)
    -> (trees: list[MerkleTree], final_poly_coeffs (n,2))"""
    coeffs = np.ascontiguousarray(np.array(coeffs, dtype=np.uint64)).reshape(-1, 2)
    values = np.ascontiguousarray(np.array(values, dtype=np.uint64)).reshape(-1, 2)
    n = coeffs.shape[0]
    nl = len(reduction_arity_bits)
    ab = (C.c_uint32 * max(nl, 1))(*reduction_arity_bits)
    handles = (C.c_void_p * max(nl, 1))()
    n_final = (n >> sum(reduction_arity_bits)) >> rate_bits
    final = np.zeros((max(n_final, 1), 2), np.uint64)
    ctx.check(ctx.lib.p2b_fri_commit(ctx.h, _ptr(coeffs), _ptr(values), n, ab, nl, rate_bits, cap_height,
                                     challenger.h, handles, _ptr(final)))
    trees = []
    ln = n
    for i, a in enumerate(reduction_arity_bits):
        t = MerkleTree(ctx, C.c_void_p(handles[i]))
        t.leaf_len = 2 << a
        trees.append(t)
        ln >>= a
    return trees, final[:n_final]


def fri_proof_of_work(ctx, challenger, proof_of_work_bits):
    """fri::prover::fri_proof_of_work(challenger, config) -> pow_witness (the minimal one)"""
    w = C.c_uint64()
    ctx.check(ctx.lib.p2b_fri_pow(ctx.h, challenger.h, proof_of_work_bits, C.byref(w)))
    return int(w.value)


class CircuitData:
    """The slice of plonky2's CommonCircuitData the prover stages between the commitments read (gates +
    selectors_info, wire/constant counts, num_challenges, quotient_degree_factor, num_partial_products, k_is;
    cf. the dump at city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145), uploaded once per circuit.

    desc: dict(degree_bits, num_wires, num_routed_wires, num_constants, num_selectors, num_challenges,
    quotient_degree_factor, num_partial_products, num_gate_constraints, k_is, gates=[dict(kind, p0, p1,
    selector_index, group_start, group_end, row)])."""

    def __init__(self, ctx, desc):
        self.ctx, self.desc = ctx, dict(desc)
        gates = (_lib.GateStruct * len(desc["gates"]))(*[
            _lib.GateStruct(g["kind"], g.get("p0", 0), g.get("p1", 0), g["selector_index"], g["group_start"],
                            g["group_end"], g["row"]) for g in desc["gates"]])
        k_is = np.ascontiguousarray(np.array(desc["k_is"], dtype=np.uint64))
        d = _lib.CircuitDescStruct(desc["degree_bits"], desc["num_wires"], desc["num_routed_wires"],
                                   desc["num_constants"], desc["num_selectors"], desc["num_challenges"],
                                   desc["quotient_degree_factor"], desc["num_partial_products"],
                                   desc["num_gate_constraints"], len(desc["gates"]), gates, _ptr(k_is))
        h = C.c_void_p()
        ctx.check(ctx.lib.p2b_circuit_new(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def free(self):
        if self.h:
            self.ctx.lib.p2b_circuit_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def _felts(x):
    return np.ascontiguousarray(np.array([int(v) for v in x], dtype=np.uint64))


def all_wires_permutation_partial_products(ctx, circuit, constants_sigmas_commitment, wires_commitment, betas,
                                           gammas, rate_bits, cap_height):
    """plonk::prover::all_wires_permutation_partial_products + the from_values commit of
    [Zs, partial products] that prove_with_partition_witness performs next -> PolynomialBatch"""
    h = C.c_void_p()
    ctx.check(ctx.lib.p2b_zs_partial_products_commit(ctx.h, circuit.h, constants_sigmas_commitment.h,
                                                     wires_commitment.h, _ptr(_felts(betas)), _ptr(_felts(gammas)),
                                                     rate_bits, cap_height, C.byref(h)))
    return PolynomialBatch(ctx, h)


def compute_quotient_polys(ctx, circuit, constants_sigmas_commitment, public_inputs_hash, wires_commitment,
                           zs_partial_products_commitment, betas, gammas, alphas, rate_bits, cap_height):
    """plonk::prover::compute_quotient_polys + chunking + from_coeffs -> quotient_polys_commitment
    (PolynomialBatch of num_challenges * quotient_degree_factor polynomials)"""
    h = C.c_void_p()
    ctx.check(ctx.lib.p2b_quotient_commit(ctx.h, circuit.h, constants_sigmas_commitment.h, wires_commitment.h,
                                          zs_partial_products_commitment.h, _ptr(_felts(public_inputs_hash)),
                                          _ptr(_felts(betas)), _ptr(_felts(gammas)), _ptr(_felts(alphas)), rate_bits,
                                          cap_height, C.byref(h)))
    return PolynomialBatch(ctx, h)


class FriParams:
    """plonky2 FriParams / FriConfig as the reference serialises them
    (city_common_circuit/src/verify_template/ser_data.rs:56-154)"""

    def __init__(self, rate_bits=3, cap_height=4, proof_of_work_bits=16, num_query_rounds=28,
                 reduction_arity_bits=(4, 4)):
        self.rate_bits, self.cap_height = rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.reduction_arity_bits = list(reduction_arity_bits)

    def struct(self):
        s = _lib.FriParamsStruct(self.rate_bits, self.cap_height, self.proof_of_work_bits, self.num_query_rounds,
                                 len(self.reduction_arity_bits))
        for i, a in enumerate(self.reduction_arity_bits):
            s.reduction_arity_bits[i] = a
        return s


def prove_openings(ctx, instance_batches, oracles, challenger, fri_params):
    """PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing) -> FriProof as a dict.
    instance_batches: [(point (2,), [(oracle_index, first, count), ...]), ...] (FriInstanceInfo.batches)."""
    nb = len(instance_batches)
    fb = (_lib.FriBatchStruct * nb)()
    for i, (point, ranges) in enumerate(instance_batches):
        fb[i].point[0], fb[i].point[1] = int(point[0]), int(point[1])
        fb[i].n_ranges = len(ranges)
        for j, (o, first, count) in enumerate(ranges):
            fb[i].ranges[j] = _lib.FriRange(o, first, count)
    handles = (C.c_void_p * len(oracles))(*[o.h for o in oracles])
    ps = fri_params.struct()
    n_words = int(ctx.lib.p2b_fri_proof_len(handles, len(oracles), C.byref(ps)))
    if n_words == 0:
        raise P2BError(-1, "inconsistent FRI parameters")
    buf = np.zeros(n_words, np.uint64)
    ctx.check(ctx.lib.p2b_prove_openings(ctx.h, handles, len(oracles), fb, nb, challenger.h, C.byref(ps), _ptr(buf),
                                         n_words))
    widths = [(o.n_cols, o.merkle_tree.cap_height) for o in oracles]
    proof, pos = _parse_fri_proof(buf, 0, widths, oracles[0].degree_log, fri_params)
    assert pos == n_words
    return proof


def _parse_fri_proof(buf, pos, widths, degree_log, fri_params):
    """split the flat words p2b_prove_openings writes (FriProof's field order); widths = [(n_cols, cap_height)]"""
    log_N = degree_log + fri_params.rate_bits
    cap_words = 4 << fri_params.cap_height
    caps = []
    for _ in fri_params.reduction_arity_bits:
        caps.append(buf[pos:pos + cap_words].reshape(-1, 4))
        pos += cap_words
    rounds = []
    for _ in range(fri_params.num_query_rounds):
        initial = []
        for n_cols, cap_h in widths:
            L = log_N - cap_h
            leaf = buf[pos:pos + n_cols]
            pos += n_cols
            sib = buf[pos:pos + 4 * L].reshape(-1, 4)
            pos += 4 * L
            initial.append((leaf, sib))
        steps = []
        lc = log_N
        for a in fri_params.reduction_arity_bits:
            lc -= a
            ev = buf[pos:pos + (2 << a)].reshape(-1, 2)
            pos += 2 << a
            L = lc - fri_params.cap_height
            sib = buf[pos:pos + 4 * L].reshape(-1, 4)
            pos += 4 * L
            steps.append((ev, sib))
        rounds.append(dict(initial_trees_proof=initial, steps=steps))
    n_final = ((1 << log_N) >> sum(fri_params.reduction_arity_bits)) >> fri_params.rate_bits
    final_poly = buf[pos:pos + 2 * n_final].reshape(-1, 2)
    pos += 2 * n_final
    pow_witness = int(buf[pos])
    pos += 1
    return dict(commit_phase_merkle_caps=caps, query_round_proofs=rounds, final_poly=final_poly,
                pow_witness=pow_witness), pos


def prove_native(ctx, circuit, constants_sigmas_commitment, circuit_digest, wire_values, public_inputs, fri_params,
                 raw=False):
    """The same flow as prove() in ONE library call (p2b_prove): the transcript never leaves the device and the
    host synchronises only for the proof-of-work search and the final download.  This is the call a patched
    `CircuitData::prove` makes after witness generation."""
    d = circuit.desc
    keep, ptrs, log_n, n_cols = PolynomialBatch._cols(wire_values)
    if n_cols != d["num_wires"] or log_n != d["degree_bits"]:
        raise ValueError("witness shape does not match the circuit")
    pis = _felts(public_inputs) if len(public_inputs) else np.zeros(1, np.uint64)
    ps = fri_params.struct()
    cs = constants_sigmas_commitment
    n_words = int(ctx.lib.p2b_proof_len(circuit.h, cs.h, C.byref(ps), len(public_inputs)))
    if n_words == 0:
        raise P2BError(-1, "inconsistent FRI parameters")
    buf = np.zeros(n_words, np.uint64)
    ctx.check(ctx.lib.p2b_prove(ctx.h, circuit.h, cs.h, _ptr(_felts(circuit_digest)), ptrs, _ptr(pis),
                                len(public_inputs), C.byref(ps), _ptr(buf), n_words))
    if raw:
        return buf
    return parse_proof_words(circuit.desc, cs, fri_params, buf, len(public_inputs))


def circuit_fingerprint(ctx, constants_sigmas_commitment, circuit_digest):
    """get_circuit_fingerprint_generic(verifier_data) = hash_no_pad(constants_sigmas_cap ‖ circuit_digest)
    (city_common_circuit/src/proof_minifier/pm_core.rs:18-42; the in-circuit twin: builder/verify.rs:41-53): how City
    Rollup names a circuit — the leaves of the sighash whitelist tree and the `allowed_fingerprints` of the aggregators.
    Computed on the device from the batch's cap."""
    cap = np.asarray(constants_sigmas_commitment.cap, dtype=np.uint64).reshape(-1)
    return ctx.hash_no_pad(np.concatenate([cap, _felts(circuit_digest)]))


def prove_submit(ctx, circuit, constants_sigmas_commitment, circuit_digest, wire_values, public_inputs, fri_params, wait_upload=True):
    """p2b_prove_submit: enqueue the whole proof and return; `wire_values` may be reused at once.  -> number of proof
    words to hand to prove_collect.  One proof may be pending per context.  wait_upload=False is
    p2b_prove_submit_nowait: pinned `wire_values` must then stay untouched until prove_upload_poll(ctx) is True."""
    keep, ptrs, log_n, n_cols = PolynomialBatch._cols(wire_values)
    d = circuit.desc
    if n_cols != d["num_wires"] or log_n != d["degree_bits"]:
        raise ValueError("witness shape does not match the circuit")
    pis = _felts(public_inputs) if len(public_inputs) else np.zeros(1, np.uint64)
    ps = fri_params.struct()
    cs = constants_sigmas_commitment
    n_words = int(ctx.lib.p2b_proof_len(circuit.h, cs.h, C.byref(ps), len(public_inputs)))
    if n_words == 0:
        raise P2BError(-1, "inconsistent FRI parameters")
    fn = ctx.lib.p2b_prove_submit if wait_upload else ctx.lib.p2b_prove_submit_nowait
    ctx.check(fn(ctx.h, circuit.h, cs.h, _ptr(_felts(circuit_digest)), ptrs, _ptr(pis), len(public_inputs), C.byref(ps)))
    if not wait_upload:
        ctx._upload_keep = keep  # the DMA still reads these
    return n_words


def prove_upload_poll(ctx):
    """True once the witness of the submitted proof has been read (p2b_prove_upload_poll)"""
    rc = ctx.lib.p2b_prove_upload_poll(ctx.h)
    if rc < 0:
        ctx.check(rc)
    return rc == 1


def prove_poll(ctx):
    rc = ctx.lib.p2b_prove_poll(ctx.h)
    if rc < 0:
        ctx.check(rc)
    return rc == 1


def prove_collect(ctx, n_words):
    buf = np.zeros(n_words, np.uint64)
    ctx.check(ctx.lib.p2b_prove_collect(ctx.h, _ptr(buf), n_words))
    return buf


def prove_native_device(ctx, circuit, constants_sigmas_commitment, circuit_digest, wire_values_dev_ptr, public_inputs,
                        fri_params):
    """p2b_prove_dev: the witness (num_wires x 2^degree_bits u64, column-major) already lives in HBM -> proof words"""
    pis = _felts(public_inputs) if len(public_inputs) else np.zeros(1, np.uint64)
    ps = fri_params.struct()
    cs = constants_sigmas_commitment
    n_words = int(ctx.lib.p2b_proof_len(circuit.h, cs.h, C.byref(ps), len(public_inputs)))
    if n_words == 0:
        raise P2BError(-1, "inconsistent FRI parameters")
    buf = np.zeros(n_words, np.uint64)
    ctx.check(ctx.lib.p2b_prove_dev(ctx.h, circuit.h, cs.h, _ptr(_felts(circuit_digest)), C.c_void_p(wire_values_dev_ptr),
                                    _ptr(pis), len(public_inputs), C.byref(ps), _ptr(buf), n_words))
    return buf


def parse_proof_words(d, cs, fri_params, buf, n_public_inputs):
    """flat words (p2b_prove's output) -> proof dict in ProofWithPublicInputs' field names"""
    n_words = buf.size
    nch, nc, nr = d["num_challenges"], d["num_constants"], d["num_routed_wires"]
    npp, qdf = d["num_partial_products"], d["quotient_degree_factor"]
    cap_words = 4 << fri_params.cap_height
    pos = 0

    def take(n_words_, shape):
        nonlocal pos
        a = buf[pos:pos + n_words_].reshape(shape)
        pos += n_words_
        return a

    proof = dict(wires_cap=take(cap_words, (-1, 4)), plonk_zs_partial_products_cap=take(cap_words, (-1, 4)),
                 quotient_polys_cap=take(cap_words, (-1, 4)))
    op = {}
    for k, cnt in (("constants", nc), ("plonk_sigmas", nr), ("wires", d["num_wires"]), ("plonk_zs", nch),
                   ("plonk_zs_next", nch), ("partial_products", nch * npp), ("quotient_polys", nch * qdf)):
        op[k] = take(2 * cnt, (-1, 2))
    proof["openings"] = op
    widths = [(cs.n_cols, cs.merkle_tree.cap_height), (d["num_wires"], fri_params.cap_height),
              (nch * (1 + npp), fri_params.cap_height), (nch * qdf, fri_params.cap_height)]
    proof["opening_proof"], pos = _parse_fri_proof(buf, pos, widths, d["degree_bits"], fri_params)
    proof["public_inputs"] = [int(x) for x in buf[pos:pos + n_public_inputs]]
    assert pos + n_public_inputs == n_words
    return proof


def prove(ctx, circuit, constants_sigmas_commitment, circuit_digest, wire_values, public_inputs, fri_params):
    """plonk::prover::prove_with_partition_witness from the filled witness onwards (witness generation is the
    caller's: SURVEY.md §8 scope), for circuits without lookups / blinding:
    wires commit -> betas, gammas -> Z / partial products commit -> alphas -> quotient commit -> zeta ->
    openings -> FRI proof.  Returns ProofWithPublicInputs as a dict (all field elements canonical)."""
    d = circuit.desc
    nch, rb, ch_ = d["num_challenges"], fri_params.rate_bits, fri_params.cap_height
    public_inputs = [int(x) for x in public_inputs]
    public_inputs_hash = ctx.hash_no_pad(public_inputs)
    wires = PolynomialBatch.from_values(ctx, wire_values, rb, False, ch_, keep_values=True)
    challenger = Challenger(ctx)
    challenger.observe_hash(circuit_digest)
    challenger.observe_hash(public_inputs_hash)
    challenger.observe_cap(wires)
    betas = challenger.get_n_challenges(nch)
    gammas = challenger.get_n_challenges(nch)
    zs_pp = all_wires_permutation_partial_products(ctx, circuit, constants_sigmas_commitment, wires, betas, gammas,
                                                   rb, ch_)
    challenger.observe_cap(zs_pp)
    alphas = challenger.get_n_challenges(nch)
    quotient = compute_quotient_polys(ctx, circuit, constants_sigmas_commitment, public_inputs_hash, wires, zs_pp,
                                      betas, gammas, alphas, rb, ch_)
    challenger.observe_cap(quotient)
    zeta = challenger.get_extension_challenge()
    n = 1 << d["degree_bits"]
    # plonky2: ensure!(zeta.exp_power_of_2(degree_bits) != F::Extension::ONE, "Opening point is in the subgroup.")
    z0, z1 = int(zeta[0]) % P, int(zeta[1]) % P
    for _ in range(d["degree_bits"]):
        z0, z1 = (z0 * z0 + 7 * z1 * z1) % P, (2 * z0 * z1) % P
    if (z0, z1) == (1, 0):
        raise P2BError(-1, "Opening point is in the subgroup.")
    # g = primitive n-th root
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - d["degree_bits"]), P) if d["degree_bits"] else 1
    zeta_next = [zeta[0] * g % P, zeta[1] * g % P]
    cs = constants_sigmas_commitment
    nc = d["num_constants"]
    cs_z = cs.eval_ext(zeta)
    openings = dict(constants=cs_z[:nc], plonk_sigmas=cs_z[nc:], wires=wires.eval_ext(zeta))
    zs_z = zs_pp.eval_ext(zeta)
    openings["plonk_zs"], openings["partial_products"] = zs_z[:nch], zs_z[nch:]
    openings["plonk_zs_next"] = zs_pp.eval_ext(zeta_next, 0, nch)
    openings["quotient_polys"] = quotient.eval_ext(zeta)
    # challenger.observe_openings(&openings.to_fri_openings()): the zeta batch, then the zeta_next batch
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "plonk_zs_next"):
        challenger.observe_extension_elements(openings[k])
    oracles = [cs, wires, zs_pp, quotient]
    instance = [(zeta, [(o, 0, oracles[o].n_cols) for o in range(4)]), (zeta_next, [(2, 0, nch)])]
    opening_proof = prove_openings(ctx, instance, oracles, challenger, fri_params)
    proof = dict(wires_cap=wires.cap, plonk_zs_partial_products_cap=zs_pp.cap, quotient_polys_cap=quotient.cap,
                 openings=openings, opening_proof=opening_proof, public_inputs=public_inputs)
    for b in (quotient, zs_pp, wires):
        b.free()
    challenger.free()
    return proof


P = 0xFFFFFFFF00000001


# ----------------------------------------------------------------------------- proof bytes (bincode)
def proof_shape(desc, n_public_inputs, constants_sigmas_cap_height):
    """p2b_proof_shape from a CommonCircuitData-shaped dict (CircuitData.desc / tests/golden/circuit_params.json)"""
    return _lib.ProofShapeStruct(desc["degree_bits"], desc["num_constants"], desc["num_routed_wires"],
                                 desc["num_wires"], desc["num_challenges"], desc["num_partial_products"],
                                 desc["quotient_degree_factor"], constants_sigmas_cap_height, n_public_inputs)


def proof_to_bincode(shape, fri_params, words):
    """`bincode::serialize(&proof_with_pis)` as the reference's proof store does it
    (city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46): the u64 words of prove_native(raw=True)
    -> the bytes the store holds."""
    lib = _lib.load()
    ps = fri_params.struct()
    words = np.ascontiguousarray(words, dtype=np.uint64)
    n = int(lib.p2b_proof_bincode_len(C.byref(shape), C.byref(ps)))
    if n == 0:
        raise P2BError(-1, "inconsistent proof shape / FRI parameters")
    out = (C.c_uint8 * n)()
    written = C.c_size_t()
    rc = lib.p2b_proof_to_bincode(C.byref(shape), C.byref(ps), _ptr(words), len(words), out, n, C.byref(written))
    if rc != 0:
        raise P2BError(rc, "proof words do not match the shape (%d words)" % len(words))
    return bytes(out[:written.value])


def proof_from_bincode(shape, fri_params, blob):
    """`bincode::deserialize::<ProofWithPublicInputs<F, C, D>>` (memory_proof_store/mod.rs:65-72) -> the u64 words
    in p2b_prove's layout; raises P2BError when any length prefix disagrees with the shape."""
    lib = _lib.load()
    ps = fri_params.struct()
    n = int(lib.p2b_proof_words(C.byref(shape), C.byref(ps)))
    if n == 0:
        raise P2BError(-1, "inconsistent proof shape / FRI parameters")
    words = np.zeros(n, np.uint64)
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob) if len(blob) else (C.c_uint8 * 1)()
    got = C.c_size_t()
    rc = lib.p2b_proof_from_bincode(C.byref(shape), C.byref(ps), buf, len(blob), _ptr(words), n, C.byref(got))
    if rc != 0:
        raise P2BError(rc, "blob is not a proof of this shape")
    return words
