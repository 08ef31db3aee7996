"""ctypes binding of libp2b.so (include/p2b.h).  Loading fails loudly when the CUDA library has not
been built; there is no CPU fallback (the oracle under oracle/ is test infrastructure only)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("P2B_LIB") or os.path.join(_HERE, "libp2b.so")  # P2B_LIB: an alternative build (kernel tuning)

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)
u32 = C.c_uint32
vp = C.c_void_p
sz = C.c_size_t

# name -> (restype, argtypes).  Every symbol include/p2b.h declares is listed here
# (tests/test_abi.py checks the header against this table and against the built library).
SIGNATURES = {
    "p2b_version": (C.c_int, []),
    "p2b_init": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "p2b_init_on_stream": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "p2b_destroy": (None, [vp]),
    "p2b_last_error": (C.c_char_p, [vp]),
    "p2b_synchronize": (C.c_int, [vp]),
    "p2b_set_blocking_sync": (C.c_int, [vp, C.c_int]),
    "p2b_set_latency_mode": (C.c_int, [vp, C.c_int]),
    "p2b_host_alloc": (C.c_int, [vp, sz, C.POINTER(vp)]),
    "p2b_host_free": (C.c_int, [vp, vp]),
    "p2b_launch_count": (u64, [vp]),
    "p2b_timer_start": (C.c_int, [vp]),
    "p2b_timer_stop_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "p2b_timer_span_ms": (C.c_int, [vp, vp, C.POINTER(C.c_float)]),
    "p2b_profile_enable": (C.c_int, [vp, C.c_int]),
    "p2b_profile_read": (C.c_int, [vp, C.POINTER(C.c_float), u64p]),
    "p2b_batch_from_values": (C.c_int, [vp, C.POINTER(u64p), sz, u32, u32, u32, u32, C.POINTER(vp)]),
    "p2b_batch_from_coeffs": (C.c_int, [vp, C.POINTER(u64p), sz, u32, u32, u32, u32, C.POINTER(vp)]),
    "p2b_batch_from_values_dev": (C.c_int, [vp, vp, sz, u32, u32, u32, u32, C.POINTER(vp)]),
    "p2b_batch_from_coeffs_dev": (C.c_int, [vp, vp, sz, u32, u32, u32, u32, C.POINTER(vp)]),
    "p2b_batch_free": (None, [vp]),
    "p2b_batch_attach": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "p2b_batch_export_len": (sz, [vp]),
    "p2b_batch_export": (C.c_int, [vp, C.POINTER(C.c_uint8), sz, C.POINTER(sz)]),
    "p2b_batch_import": (C.c_int, [vp, C.POINTER(C.c_uint8), sz, C.POINTER(vp)]),
    "p2b_batch_n_cols": (sz, [vp]),
    "p2b_batch_degree_log": (u32, [vp]),
    "p2b_batch_rate_bits": (u32, [vp]),
    "p2b_batch_tree": (vp, [vp]),
    "p2b_batch_cap": (C.c_int, [vp, u64p]),
    "p2b_batch_coeffs": (C.c_int, [vp, sz, u64p]),
    "p2b_batch_leaf": (C.c_int, [vp, sz, u64p]),
    "p2b_batch_lde_values": (C.c_int, [vp, sz, sz, u64p]),
    "p2b_batch_leaves": (C.c_int, [vp, u64p]),
    "p2b_batch_dev_lde": (vp, [vp]),
    "p2b_batch_dev_coeffs": (vp, [vp]),
    "p2b_batch_values": (C.c_int, [vp, sz, u64p]),
    "p2b_batch_lde_col": (C.c_int, [vp, sz, u64p]),
    "p2b_circuit_new": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "p2b_circuit_free": (None, [vp]),
    "p2b_zs_partial_products_commit": (C.c_int, [vp, vp, vp, vp, u64p, u64p, u32, u32, C.POINTER(vp)]),
    "p2b_quotient_commit": (C.c_int, [vp, vp, vp, vp, vp, u64p, u64p, u64p, u64p, u32, u32, C.POINTER(vp)]),
    "p2b_merkle_new": (C.c_int, [vp, u64p, sz, sz, u32, C.POINTER(vp)]),
    "p2b_tree_free": (None, [vp]),
    "p2b_tree_n_leaves": (sz, [vp]),
    "p2b_tree_cap_height": (u32, [vp]),
    "p2b_tree_cap": (C.c_int, [vp, u64p]),
    "p2b_tree_prove": (C.c_int, [vp, sz, u64p]),
    "p2b_tree_digests": (C.c_int, [vp, u64p]),
    "p2b_tree_leaf": (C.c_int, [vp, sz, u64p]),
    "p2b_poseidon_permute": (C.c_int, [vp, u64p, sz]),
    "p2b_hash_no_pad": (C.c_int, [vp, u64p, sz, u64p]),
    "p2b_two_to_one": (C.c_int, [vp, u64p, u64p, sz, u64p]),
    "p2b_challenger_new": (C.c_int, [vp, C.POINTER(vp)]),
    "p2b_challenger_free": (None, [vp]),
    "p2b_challenger_observe": (C.c_int, [vp, u64p, sz]),
    "p2b_challenger_observe_cap": (C.c_int, [vp, vp]),
    "p2b_challenger_get": (C.c_int, [vp, sz, u64p]),
    "p2b_challenger_export": (C.c_int, [vp, u64p]),
    "p2b_challenger_import": (C.c_int, [vp, u64p]),
    "p2b_fri_commit": (C.c_int, [vp, u64p, u64p, sz, C.POINTER(u32), sz, u32, u32, vp, C.POINTER(vp), u64p]),
    "p2b_fri_pow": (C.c_int, [vp, vp, u32, u64p]),
    "p2b_batch_eval_ext": (C.c_int, [vp, u64p, sz, sz, u64p]),
    "p2b_fri_proof_len": (sz, [C.POINTER(vp), sz, vp]),
    "p2b_prove_openings": (C.c_int, [vp, C.POINTER(vp), sz, vp, sz, vp, vp, u64p, sz]),
    "p2b_proof_len": (sz, [vp, vp, vp, sz]),
    "p2b_prove": (C.c_int, [vp, vp, vp, u64p, C.POINTER(u64p), u64p, sz, vp, u64p, sz]),
    "p2b_prove_dev": (C.c_int, [vp, vp, vp, u64p, vp, u64p, sz, vp, u64p, sz]),
    "p2b_prove_submit": (C.c_int, [vp, vp, vp, u64p, C.POINTER(u64p), u64p, sz, vp]),
    "p2b_prove_poll": (C.c_int, [vp]),
    "p2b_prove_submit_nowait": (C.c_int, [vp, vp, vp, u64p, C.POINTER(u64p), u64p, sz, vp]),
    "p2b_prove_upload_poll": (C.c_int, [vp]),
    "p2b_prove_collect": (C.c_int, [vp, u64p, sz]),
    "p2b_plan_info": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
    "p2b_proof_words": (sz, [vp, vp]),
    "p2b_proof_bincode_len": (sz, [vp, vp]),
    "p2b_proof_to_bincode": (C.c_int, [vp, vp, u64p, sz, C.POINTER(C.c_uint8), sz, C.POINTER(sz)]),
    "p2b_proof_from_bincode": (C.c_int, [vp, vp, C.POINTER(C.c_uint8), sz, u64p, sz, C.POINTER(sz)]),
}



class GateStruct(C.Structure):  # p2b_gate
    _fields_ = [("kind", u32), ("p0", u32), ("p1", u32), ("selector_index", u32), ("group_start", u32),
                ("group_end", u32), ("row", u32)]


class CircuitDescStruct(C.Structure):  # p2b_circuit_desc
    _fields_ = [("degree_bits", u32), ("num_wires", u32), ("num_routed_wires", u32), ("num_constants", u32),
                ("num_selectors", u32), ("num_challenges", u32), ("quotient_degree_factor", u32),
                ("num_partial_products", u32), ("num_gate_constraints", u32), ("n_gates", u32),
                ("gates", C.POINTER(GateStruct)), ("k_is", u64p)]


class ProofShapeStruct(C.Structure):  # p2b_proof_shape
    _fields_ = [("degree_bits", u32), ("num_constants", u32), ("num_routed_wires", u32), ("num_wires", u32),
                ("num_challenges", u32), ("num_partial_products", u32), ("quotient_degree_factor", u32),
                ("constants_sigmas_cap_height", u32), ("n_public_inputs", u32)]


class FriRange(C.Structure):
    _fields_ = [("oracle", u32), ("first", u32), ("count", u32)]


class FriBatchStruct(C.Structure):  # p2b_fri_batch
    _fields_ = [("point", u64 * 2), ("n_ranges", u32), ("ranges", FriRange * 8)]


class FriParamsStruct(C.Structure):  # p2b_fri_params
    _fields_ = [("rate_bits", u32), ("cap_height", u32), ("proof_of_work_bits", u32), ("num_query_rounds", u32),
                ("n_layers", u32), ("reduction_arity_bits", u32 * 16)]


_lib = None


def build(verbose=False):
    """Compile csrc/ for sm_100a into city_rollup_b200/libp2b.so (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc")]
    if not verbose:
        args.append("-s")
    subprocess.check_call(args)
    return SO_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA path)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
