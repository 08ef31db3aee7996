"""The proof byte format (`bincode::serialize(&ProofWithPublicInputs)`, the form the reference's proof store holds:
city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46,65-72) behind p2b_proof_to_bincode /
p2b_proof_from_bincode, pinned on the ten real proofs inside qbench_data/example.bin (tests/golden/): every stored blob
must survive blob -> words -> blob byte for byte, and the words must be the fields the independent Python parser
(tests/proof_parser.py) reads.  Host-only formatting code: runs without a GPU."""
import json
import os

import numpy as np
import pytest

import city_rollup_b200 as m
from proof_parser import parse_proof


@pytest.fixture(scope="module")
def params(golden_dir):
    return json.load(open(os.path.join(golden_dir, "circuit_params.json")))


@pytest.fixture(scope="module")
def blobs(golden_dir):
    idx = json.load(open(os.path.join(golden_dir, "example_proofs.json")))["proofs"]
    blob = open(os.path.join(golden_dir, "example_proofs.bin"), "rb").read()
    return [blob[e["offset"]: e["offset"] + e["len"]] for e in idx]


def _fri(params):
    return m.FriParams(params["rate_bits"], params["cap_height"], params["proof_of_work_bits"], params["num_query_rounds"],
                       params["reduction_arity_bits"])


def flatten(p):
    """the u64 words of a parsed proof in p2b_prove's order (include/p2b.h)"""
    out = [p["wires_cap"], p["zs_pp_cap"], p["quotient_cap"]]
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys"):
        out.append(p["openings"][k])
    out += p["commit_phase_merkle_caps"]
    for r in p["query_rounds"]:
        for leaf, sib in r["initial"]:
            out += [leaf, sib]
        for ev, sib in r["steps"]:
            out += [ev, sib]
    out += [p["final_poly"], np.array([p["pow_witness"]], np.uint64), p["public_inputs"]]
    return np.concatenate([np.asarray(a, np.uint64).reshape(-1) for a in out])


def test_stored_proofs_round_trip_byte_for_byte(params, blobs):
    fp = _fri(params)
    assert len(blobs) == 10
    for blob in blobs:
        parsed = parse_proof(blob)
        shape = m.proof_shape(params, len(parsed["public_inputs"]), params["cap_height"])
        words = m.proof_from_bincode(shape, fp, blob)
        assert (words == flatten(parsed)).all()
        assert m.proof_to_bincode(shape, fp, words) == blob


def test_wrong_shape_is_rejected(params, blobs):
    fp = _fri(params)
    blob = blobs[0]
    npi = len(parse_proof(blob)["public_inputs"])
    good = m.proof_shape(params, npi, params["cap_height"])
    words = m.proof_from_bincode(good, fp, blob)
    # a blob of another circuit: one more wire moves 2 opening words + 28 leaf words (length mismatch)
    other = dict(params, num_wires=params["num_wires"] + 1)
    with pytest.raises(m.P2BError):
        m.proof_from_bincode(m.proof_shape(other, npi, params["cap_height"]), fp, blob)
    # same total length, different split: one wire more, one routed wire less in the openings / leaves
    swapped = dict(params, num_wires=params["num_wires"] + 1, num_constants=params["num_constants"] - 1)
    with pytest.raises(m.P2BError):
        m.proof_from_bincode(m.proof_shape(swapped, npi, params["cap_height"]), fp, blob)
    with pytest.raises(m.P2BError):  # truncated
        m.proof_from_bincode(good, fp, blob[:-8])
    corrupt = bytearray(blob)
    corrupt[0] ^= 1  # the first length prefix (wires_cap: 16 digests)
    with pytest.raises(m.P2BError):
        m.proof_from_bincode(good, fp, bytes(corrupt))
    with pytest.raises(m.P2BError):  # word count must match the shape
        m.proof_to_bincode(good, fp, words[:-1])
    with pytest.raises(m.P2BError):  # sum of arities above degree_bits
        m.proof_to_bincode(good, m.FriParams(3, 4, 16, 28, [4, 4, 5]), words)


def test_lengths(params):
    fp = _fri(params)
    shape = m.proof_shape(params, 4, params["cap_height"])
    lib = m.load()
    import ctypes as C

    ps = fp.struct()
    lib.p2b_proof_words.restype = C.c_size_t
    n_words = lib.p2b_proof_words(C.byref(shape), C.byref(ps))
    n_bytes = lib.p2b_proof_bincode_len(C.byref(shape), C.byref(ps))
    # City shape, 4 public inputs: the stored 130 360-byte blobs have 4 + ... public inputs; count prefixes instead
    n_prefix = 3 + 9 + 1 + 2 + 1 + 28 * (1 + 4 * 2 + 1 + 2 * 2) + 1 + 1
    assert n_bytes == 8 * (n_words + n_prefix)


@pytest.mark.gpu
def test_gpu_proof_bytes_parse_back(params):
    """a proof made by p2b_prove, serialised by p2b_proof_to_bincode, read by the independent parser: same fields"""
    import p2oracle as O
    import plonk_ref as R

    ctx = m.Context(0)
    gates = [(R.GATE_PUBLIC_INPUT, 0, 0), (R.GATE_NOOP, 0, 0), (R.GATE_CONSTANT, 2, 0), (R.GATE_ARITHMETIC, 20, 0),
             (R.GATE_POSEIDON, 0, 0)]
    pis = [11, 22, 33]
    circ = R.SyntheticCircuit(6, gates, [(0, 4), (4, 5)], 3, pi_hash=O.hash_no_pad(pis))
    fp = m.FriParams(3, 2, 8, 5, [3, 2])
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 2, keep_values=True)
    words = m.prove_native(ctx, cd, cs, [5, 6, 7, 8], circ.wire_values(), pis, fp, raw=True)
    shape = m.proof_shape(cd.desc, len(pis), 2)
    blob = m.proof_to_bincode(shape, fp, words)
    parsed = parse_proof(blob)
    assert (flatten(parsed) == words).all()
    assert [int(x) for x in parsed["public_inputs"]] == pis
    assert len(parsed["query_rounds"]) == 5 and len(parsed["commit_phase_merkle_caps"]) == 2
    assert len(parsed["openings"]["lookup_zs"]) == 0 and len(parsed["openings"]["lookup_zs_next"]) == 0
    assert (m.proof_from_bincode(shape, fp, blob) == words).all()
    cs.free()
    cd.free()
    ctx.close()
