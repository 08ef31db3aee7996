"""ctypes/numpy front-end of the CPU oracle (oracle/libp2oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (city_rollup_b200) never imports it.
Function-by-function provenance is in oracle/p2oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libp2oracle.so")
P = 0xFFFFFFFF00000001

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)


def build(force=False):
    # `make` is a no-op when the .so is newer than the sources; on the GPU box the prebuilt .so
    # travels with the snapshot and make only checks timestamps.
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(_SO):
            raise
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.gl_canon.restype = u64
        L.gl_canon.argtypes = [u64]
        for f in ("gl_add", "gl_sub", "gl_mul", "gl_mul_slow", "gl_pow"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [u64, u64]
        L.gl_inv.restype = u64
        L.gl_inv.argtypes = [u64]
        L.gl_root_of_unity.restype = u64
        L.gl_root_of_unity.argtypes = [C.c_uint]
        L.fri_proof_of_work.restype = u64
        L.challenger_get.restype = u64
        L.p2o_num_threads.restype = C.c_uint
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def arr(x):
    return np.ascontiguousarray(np.array(x, dtype=np.uint64))


# ---- field ----
def mul(a, b):
    return lib().gl_mul(int(a), int(b))


def add(a, b):
    return lib().gl_add(int(a), int(b))


def sub(a, b):
    return lib().gl_sub(int(a), int(b))


def inv(a):
    return lib().gl_inv(int(a))


def gpow(a, e):
    return lib().gl_pow(int(a), int(e))


def root_of_unity(log_n):
    return lib().gl_root_of_unity(log_n)


def ext_mul(a, b):
    o = np.zeros(2, np.uint64)
    lib().gl2_mul(_p(arr(a)), _p(arr(b)), _p(o))
    return [int(o[0]), int(o[1])]


def ext_inv(a):
    o = np.zeros(2, np.uint64)
    lib().gl2_inv(_p(arr(a)), _p(o))
    return [int(o[0]), int(o[1])]


# ---- poseidon ----
def permute(state):
    s = arr(state).copy()
    assert s.shape == (12,)
    lib().poseidon_permute(_p(s))
    return s


def hash_no_pad(x):
    x = arr(x)
    o = np.zeros(4, np.uint64)
    lib().poseidon_hash_no_pad(_p(x), C.c_size_t(x.size), _p(o))
    return o


def hash_or_noop(x):
    x = arr(x)
    o = np.zeros(4, np.uint64)
    lib().poseidon_hash_or_noop(_p(x), C.c_size_t(x.size), _p(o))
    return o


def two_to_one(l, r):
    o = np.zeros(4, np.uint64)
    lib().poseidon_two_to_one(_p(arr(l)), _p(arr(r)), _p(o))
    return o


# ---- merkle ----
def merkle_tree_new(leaves, cap_height):
    """leaves: (n_leaves, leaf_len) uint64 -> (digests (2*(n-2^cap),4) plonky2 layout, cap (2^cap,4))"""
    leaves = arr(leaves)
    n, w = leaves.shape
    ncap = 1 << cap_height
    dg = np.zeros((max(2 * (n - ncap), 1), 4), np.uint64)
    cap = np.zeros((ncap, 4), np.uint64)
    lib().merkle_tree_new(_p(leaves), C.c_size_t(n), C.c_size_t(w), C.c_uint(cap_height), _p(dg), _p(cap))
    return dg[: 2 * (n - ncap)], cap


def merkle_prove(digests, n_leaves, cap_height, index):
    nl = int(n_leaves).bit_length() - 1 - cap_height
    sib = np.zeros((max(nl, 1), 4), np.uint64)
    d = arr(digests) if len(digests) else np.zeros((1, 4), np.uint64)
    lib().merkle_prove(_p(d), C.c_size_t(n_leaves), C.c_uint(cap_height), C.c_size_t(index), _p(sib))
    return sib[:nl]


def merkle_verify(leaf, index, siblings, cap):
    leaf, sib, cap = arr(leaf), arr(siblings).reshape(-1, 4), arr(cap)
    if sib.size == 0:
        sib = np.zeros((1, 4), np.uint64)
        ns = 0
    else:
        ns = sib.shape[0]
    return bool(
        lib().merkle_verify(_p(leaf), C.c_size_t(leaf.size), C.c_size_t(index), _p(sib), C.c_uint(ns), _p(cap))
    )


# ---- fft ----
def fft(a):
    a = arr(a).copy()
    lib().gl_fft(_p(a), C.c_uint(a.size.bit_length() - 1))
    return a


def ifft(a):
    a = arr(a).copy()
    lib().gl_ifft(_p(a), C.c_uint(a.size.bit_length() - 1))
    return a


def coset_fft(a, shift=7):
    a = arr(a).copy()
    lib().gl_coset_fft(_p(a), C.c_uint(a.size.bit_length() - 1), u64(shift))
    return a


def coset_ifft(a, shift=7):
    a = arr(a).copy()
    lib().gl_coset_ifft(_p(a), C.c_uint(a.size.bit_length() - 1), u64(shift))
    return a


def ext_coset_fft(a, shift=7):
    """a: (n,2) uint64"""
    a = arr(a).copy()
    lib().gl2_coset_fft(_p(a), C.c_uint(a.shape[0].bit_length() - 1), u64(shift))
    return a


# ---- PolynomialBatch ----
def _colptrs(cols):
    cols = [arr(c) for c in cols]
    ptrs = (u64p * len(cols))(*[_p(c) for c in cols])
    return cols, ptrs


def batch_from_coeffs(cols, rate_bits=3, cap_height=4, want_leaves=True, want_digests=True):
    cols, ptrs = _colptrs(cols)
    n = cols[0].size
    N = n << rate_bits
    ncap = 1 << cap_height
    leaves = np.zeros((N, len(cols)), np.uint64) if want_leaves else None
    dg = np.zeros((max(2 * (N - ncap), 1), 4), np.uint64) if want_digests else None
    cap = np.zeros((ncap, 4), np.uint64)
    lib().batch_from_coeffs(
        ptrs, C.c_size_t(len(cols)), C.c_uint(n.bit_length() - 1), C.c_uint(rate_bits), C.c_uint(cap_height),
        _p(leaves) if want_leaves else None, _p(dg) if want_digests else None, _p(cap))
    return {"leaves": leaves, "digests": dg[: 2 * (N - ncap)] if want_digests else None, "cap": cap}


def batch_from_values(cols, rate_bits=3, cap_height=4, want_leaves=True, want_digests=True):
    cols, ptrs = _colptrs(cols)
    n = cols[0].size
    N = n << rate_bits
    ncap = 1 << cap_height
    coeffs = np.zeros((len(cols), n), np.uint64)
    leaves = np.zeros((N, len(cols)), np.uint64) if want_leaves else None
    dg = np.zeros((max(2 * (N - ncap), 1), 4), np.uint64) if want_digests else None
    cap = np.zeros((ncap, 4), np.uint64)
    lib().batch_from_values(
        ptrs, C.c_size_t(len(cols)), C.c_uint(n.bit_length() - 1), C.c_uint(rate_bits), C.c_uint(cap_height),
        _p(coeffs), _p(leaves) if want_leaves else None, _p(dg) if want_digests else None, _p(cap))
    return {"coeffs": coeffs, "leaves": leaves, "digests": dg[: 2 * (N - ncap)] if want_digests else None,
            "cap": cap}


# ---- challenger / FRI ----
class ChallengerStruct(C.Structure):
    _fields_ = [("state", u64 * 12), ("inb", u64 * 8), ("out", u64 * 8), ("n_in", C.c_uint), ("n_out", C.c_uint)]


class Challenger:
    def __init__(self):
        self.c = ChallengerStruct()
        lib().challenger_init(C.byref(self.c))

    def clone(self):
        o = Challenger()
        C.memmove(C.byref(o.c), C.byref(self.c), C.sizeof(ChallengerStruct))
        return o

    def observe(self, elems):
        e = arr(elems).reshape(-1)
        if e.size:
            lib().challenger_observe(C.byref(self.c), _p(e), C.c_size_t(e.size))

    def get(self):
        return int(lib().challenger_get(C.byref(self.c)))

    def get_n(self, n):
        return [self.get() for _ in range(n)]

    def get_ext(self):
        return self.get_n(2)

    def state_words(self):
        """(state[12], input_buffer) — what the product's challenger must equal."""
        return [int(x) for x in self.c.state], [int(self.c.inb[i]) for i in range(self.c.n_in)]


def fri_committed_trees(coeffs, values, arity_bits, challenger, rate_bits=3, cap_height=4):
    """coeffs, values: (len,2) uint64.  Returns dict(caps, leaves[], digests[], final_poly, betas)."""
    coeffs, values = arr(coeffs), arr(values)
    n = coeffs.shape[0]
    ncap = 1 << cap_height
    nl = len(arity_bits)
    ab = (C.c_uint * nl)(*arity_bits)
    caps = np.zeros((nl, ncap, 4), np.uint64)
    leaves, digests = [], []
    ln = n
    for a in arity_bits:
        leaves.append(np.zeros((ln >> a, 2 << a), np.uint64))
        digests.append(np.zeros((max(2 * ((ln >> a) - ncap), 1), 4), np.uint64))
        ln >>= a
    nfinal = ln >> rate_bits
    final = np.zeros((nfinal, 2), np.uint64)
    betas = np.zeros((nl, 2), np.uint64)
    lp = (u64p * nl)(*[_p(x) for x in leaves])
    dp = (u64p * nl)(*[_p(x) for x in digests])
    lib().fri_committed_trees(_p(coeffs), _p(values), C.c_size_t(n), ab, C.c_size_t(nl), C.c_uint(rate_bits),
                              C.c_uint(cap_height), C.byref(challenger.c), _p(caps), lp, dp, _p(final), _p(betas))
    return {"caps": caps, "leaves": leaves, "digests": digests, "final_poly": final, "betas": betas}


def fri_proof_of_work(challenger, pow_bits=16):
    return int(lib().fri_proof_of_work(C.byref(challenger.c), C.c_uint(pow_bits)))


def fri_pow_check(challenger, w, pow_bits=16):
    return int(lib().fri_pow_check(C.byref(challenger.c), u64(w), C.c_uint(pow_bits)))


def fri_compute_evaluation(x, x_index_within_coset, arity_bits, evals, beta):
    ev = arr(evals).reshape(-1)
    o = np.zeros(2, np.uint64)
    lib().fri_compute_evaluation(u64(x), C.c_uint(x_index_within_coset), C.c_uint(arity_bits), _p(ev), _p(arr(beta)), _p(o))
    return [int(o[0]), int(o[1])]


def num_threads():
    return int(lib().p2o_num_threads())


def merkle_find_index(leaf, siblings, cap):
    leaf, sib, cap = arr(leaf), arr(siblings).reshape(-1, 4), arr(cap)
    lib().merkle_find_index.restype = C.c_long
    return int(lib().merkle_find_index(_p(leaf), C.c_size_t(leaf.size), _p(sib), C.c_uint(sib.shape[0]), _p(cap),
                                       C.c_size_t(cap.shape[0])))


# ---- PLONK stages between the commitments (plonk.c) ----
GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_POSEIDON, GATE_BASE_SUM = range(6)
GATE_U32_ARITHMETIC, GATE_U32_ADD_MANY, GATE_U32_SUBTRACTION, GATE_U32_RANGE_CHECK = 6, 7, 8, 9
GATE_U32_INTERLEAVE, GATE_UNINTERLEAVE_TO_U32, GATE_UNINTERLEAVE_TO_B32, GATE_COMPARISON = 10, 11, 12, 13
(GATE_ARITHMETIC_EXT, GATE_MUL_EXT, GATE_REDUCING, GATE_REDUCING_EXT, GATE_RANDOM_ACCESS,
 GATE_POSEIDON_MDS) = range(14, 20)
GATE_COSET_INTERPOLATION = 20


class GateStruct(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("p0", C.c_uint32), ("p1", C.c_uint32), ("selector_index", C.c_uint32),
                ("group_start", C.c_uint32), ("group_end", C.c_uint32), ("row", C.c_uint32)]


class CircuitStruct(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("num_wires", C.c_uint32), ("num_routed_wires", C.c_uint32),
                ("num_constants", C.c_uint32), ("num_selectors", C.c_uint32), ("num_challenges", C.c_uint32),
                ("quotient_degree_factor", C.c_uint32), ("num_partial_products", C.c_uint32),
                ("num_gate_constraints", C.c_uint32), ("n_gates", C.c_uint32),
                ("gates", C.POINTER(GateStruct)), ("k_is", u64p)]


def circuit_struct(desc):
    """desc: dict with the CircuitStruct scalar fields, 'gates' = list of dicts (GateStruct fields), 'k_is'."""
    gates = (GateStruct * len(desc["gates"]))(*[
        GateStruct(g["kind"], g.get("p0", 0), g.get("p1", 0), g["selector_index"], g["group_start"], g["group_end"],
                   g["row"]) for g in desc["gates"]])
    k_is = arr(desc["k_is"])
    cs = CircuitStruct(desc["degree_bits"], desc["num_wires"], desc["num_routed_wires"], desc["num_constants"],
                       desc["num_selectors"], desc["num_challenges"], desc["quotient_degree_factor"],
                       desc["num_partial_products"], desc["num_gate_constraints"], len(desc["gates"]),
                       gates, _p(k_is))
    cs._keep = (gates, k_is)
    return cs


def partial_products_and_zs(desc, wires, sigmas, betas, gammas):
    """wires (num_wires, n), sigmas (num_routed, n) values on H -> (num_chal * (1 + num_pp), n)"""
    cs = circuit_struct(desc)
    wires, sigmas = arr(wires), arr(sigmas)
    n = 1 << desc["degree_bits"]
    out = np.zeros((desc["num_challenges"] * (1 + desc["num_partial_products"]), n), np.uint64)
    lib().plonk_partial_products_and_zs(C.byref(cs), _p(wires), _p(sigmas), _p(arr(betas)), _p(arr(gammas)), _p(out))
    return out


def compute_quotient_polys(desc, rate_bits, cs_leaves, wires_leaves, zs_leaves, pi_hash, betas, gammas, alphas):
    """*_leaves: (n << rate_bits, width) leaf-ordered LDE rows -> (num_chal * quotient_degree_factor, n) chunks"""
    cs = circuit_struct(desc)
    n = 1 << desc["degree_bits"]
    out = np.zeros((desc["num_challenges"] * desc["quotient_degree_factor"], n), np.uint64)
    lib().plonk_compute_quotient_polys(C.byref(cs), C.c_uint(rate_bits), _p(arr(cs_leaves)), _p(arr(wires_leaves)),
                                       _p(arr(zs_leaves)), _p(arr(pi_hash)), _p(arr(betas)), _p(arr(gammas)),
                                       _p(arr(alphas)), _p(out))
    return out


def eval_gate(kind, p0, p1, wires, consts, pi_hash, max_constraints=256):
    """unfiltered constraints of one gate on one row of (canonical) values"""
    g = GateStruct(kind, p0, p1, 0, 0, 1, 0)
    out = np.zeros(max_constraints, np.uint64)
    lib().plonk_eval_gate.restype = C.c_uint
    k = lib().plonk_eval_gate(C.byref(g), _p(arr(wires)), _p(arr(consts)), _p(arr(pi_hash)), _p(out))
    return [int(x) for x in out[:k]]


# ---- the whole prover (prove.c) ----
class FriParamsStruct(C.Structure):
    _fields_ = [("rate_bits", C.c_uint32), ("cap_height", C.c_uint32), ("proof_of_work_bits", C.c_uint32),
                ("num_query_rounds", C.c_uint32), ("n_layers", C.c_uint32), ("reduction_arity_bits", C.c_uint32 * 16)]


def fri_params_struct(fp):
    """fp: dict(rate_bits, cap_height, proof_of_work_bits, num_query_rounds, reduction_arity_bits)"""
    ab = list(fp["reduction_arity_bits"])
    return FriParamsStruct(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], len(ab),
                           (C.c_uint32 * 16)(*ab))


class ProverData:
    """prover_data.constants_sigmas_commitment of a circuit, built once (not part of a proof's time)"""

    def __init__(self, desc, constants_sigmas_values, fp):
        L = lib()
        L.p2o_prover_data_new.restype = C.c_void_p
        self.desc = desc
        self.cs = circuit_struct(desc)
        self.fp = fp
        cols, ptrs = _colptrs(constants_sigmas_values)
        self.h = C.c_void_p(L.p2o_prover_data_new(C.byref(self.cs), ptrs, C.c_uint(fp["rate_bits"]), C.c_uint(fp["cap_height"])))
        self.cap = np.zeros((1 << fp["cap_height"], 4), np.uint64)
        L.p2o_prover_data_cap(self.h, _p(self.cap))

    def prove(self, circuit_digest, wire_values, public_inputs):
        """-> the flat proof words (the ones p2b_prove writes)"""
        L = lib()
        L.p2o_proof_len.restype = C.c_size_t
        L.p2o_prove.restype = C.c_size_t
        fps = fri_params_struct(self.fp)
        n_pis = len(public_inputs)
        n_words = int(L.p2o_proof_len(C.byref(self.cs), C.c_uint(self.fp["cap_height"]), C.byref(fps), C.c_size_t(n_pis)))
        assert n_words, "inconsistent FRI parameters"
        out = np.zeros(n_words, np.uint64)
        cols, ptrs = _colptrs(wire_values)
        pis = arr(public_inputs) if n_pis else np.zeros(1, np.uint64)
        got = int(L.p2o_prove(C.byref(self.cs), self.h, _p(arr(circuit_digest)), ptrs, _p(pis), C.c_size_t(n_pis),
                              C.byref(fps), _p(out), C.c_size_t(n_words)))
        assert got == n_words, "p2o_prove failed"
        return out

    def free(self):
        if self.h:
            lib().p2o_prover_data_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
