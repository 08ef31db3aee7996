"""CPU: a complete proof built from oracle primitives (tests/verifier_ref.py::oracle_prove) is accepted by the
restated plonky2 verifier, and tampered proofs are rejected — the checker the GPU prover is later held to."""
import copy

import numpy as np
import pytest

import p2oracle as O
import plonk_ref as R
import verifier_ref as V
from test_plonk_oracle import ALL_GATES, CITY_GATES, CITY_GROUPS

FP_SMALL = dict(rate_bits=3, cap_height=2, proof_of_work_bits=6, num_query_rounds=4, reduction_arity_bits=[3, 2])


def make_case(degree_bits, gates, groups, seed):
    public_inputs = [seed, 2, 3, 0xFFFFFFFF00000000]
    circ = R.SyntheticCircuit(degree_bits, gates, groups, seed, pi_hash=O.hash_no_pad(public_inputs))
    circuit_digest = [int(x) for x in O.hash_no_pad([seed, 77])]
    return circ, circuit_digest, public_inputs


@pytest.fixture(scope="module")
def case():
    circ, digest, pis = make_case(6, ALL_GATES, [(0, 4), (4, 5), (5, 8), (8, 10)], 41)
    proof, cs_cap = V.oracle_prove(circ, digest, pis, FP_SMALL)
    return circ, digest, proof, cs_cap


def test_oracle_proof_verifies(case):
    circ, digest, proof, cs_cap = case
    assert V.verify(circ, cs_cap, digest, proof, FP_SMALL)


@pytest.mark.parametrize("what", ["opening", "final_poly", "leaf", "pow", "cap", "public_input"])
def test_tampered_proof_rejected(case, what):
    circ, digest, proof, cs_cap = case
    bad = copy.deepcopy(proof)
    if what == "opening":
        bad["openings"]["wires"][3][0] ^= np.uint64(1)
    elif what == "final_poly":
        bad["opening_proof"]["final_poly"][0][1] ^= np.uint64(1)
    elif what == "leaf":
        bad["opening_proof"]["query_round_proofs"][1]["initial_trees_proof"][1][0][5] ^= np.uint64(1)
    elif what == "pow":
        bad["opening_proof"]["pow_witness"] += 1
    elif what == "cap":
        bad["quotient_polys_cap"][0][0] ^= np.uint64(1)
    else:
        bad["public_inputs"][0] += 1
    with pytest.raises(V.VerificationError):
        V.verify(circ, cs_cap, digest, bad, FP_SMALL)


@pytest.mark.parametrize("degree_bits,gates,groups,seed,fp", [
    (6, ALL_GATES, [(0, 4), (4, 5), (5, 8), (8, 10)], 41, FP_SMALL),
    (5, ALL_GATES[:5], [(0, 4), (4, 5)], 42, dict(FP_SMALL, cap_height=0, reduction_arity_bits=[1, 2, 1], num_query_rounds=3)),
    (7, ALL_GATES, [(0, 4), (4, 5), (5, 8), (8, 10)], 43, dict(FP_SMALL, cap_height=4, reduction_arity_bits=[4], proof_of_work_bits=9)),
    (6, CITY_GATES, CITY_GROUPS, 44, FP_SMALL),
])
def test_c_prover_equals_python_composition(degree_bits, gates, groups, seed, fp):
    """oracle/prove.c::p2o_prove (the C composition timed by bench.py's reference arm and compared with p2b_prove at
    the City shape) against the independent Python composition of the same primitives, word for word"""
    circ, digest, pis = make_case(degree_bits, gates, groups, seed)
    ref, ref_cap = V.oracle_prove(circ, digest, pis, fp)
    got, cap = V.oracle_prove_c(circ, digest, pis, fp)
    assert (cap == ref_cap).all()
    want = V.flatten_proof(ref)
    assert got.shape == want.shape
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, "first differing word %d of %d" % (bad[0], want.size)


def test_city_gate_set_proof_verifies():
    """a proof over the gate set of the City Rollup op circuits (all 21 gate kinds, six selector groups) from the C
    prover is accepted by the restated verifier"""
    circ, digest, pis = make_case(6, CITY_GATES, CITY_GROUPS, 45)
    words, cs_cap = V.oracle_prove_c(circ, digest, pis, FP_SMALL)
    proof = V.parse_proof(circ, FP_SMALL, words, len(pis))
    assert (V.flatten_proof(proof) == words).all()
    assert V.verify(circ, cs_cap, digest, proof, FP_SMALL)
