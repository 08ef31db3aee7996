"""Pin the CPU oracle to every known answer the reference tree holds for the hot path
(SURVEY.md §8(c) K1..K5).  CPU-only."""
import json
import os

import numpy as np
import pytest

import p2oracle as O
from proof_parser import parse_proof

P = O.P


@pytest.fixture(scope="module")
def zero_hashes(golden_dir):
    return json.load(open(os.path.join(golden_dir, "zero_hashes.json")))


@pytest.fixture(scope="module")
def params(golden_dir):
    return json.load(open(os.path.join(golden_dir, "circuit_params.json")))


@pytest.fixture(scope="module")
def proofs(golden_dir):
    idx = json.load(open(os.path.join(golden_dir, "example_proofs.json")))["proofs"]
    blob = open(os.path.join(golden_dir, "example_proofs.bin"), "rb").read()
    return [parse_proof(blob[e["offset"] : e["offset"] + e["len"]]) for e in idx]


def test_k1_two_to_one_zero_hash_chain(zero_hashes):
    """city_crypto/src/hash/cached_zero_hashes.rs:10-1036: H[0]=0, H[i]=two_to_one(H[i-1],H[i-1])"""
    t = zero_hashes["zero"]
    assert t[0] == [0, 0, 0, 0]
    for i in range(1, 128):
        assert O.two_to_one(t[i - 1], t[i - 1]).tolist() == t[i], i


def test_k2_marked_leaf_chain(zero_hashes):
    """:1039-2066: M[1]=hash_no_pad([0]*8+[1]) (two permutations, overwrite absorb), then two_to_one"""
    t = zero_hashes["marked"]
    assert O.hash_no_pad([0] * 8 + [1]).tolist() == t[1]
    for i in range(2, 128):
        assert O.two_to_one(t[i - 1], t[i - 1]).tolist() == t[i], i


def test_k4_generator_and_roots(params):
    """zk_signature2/mod.rs:58-138: k_is[i] = 7^i; and the derived 2-adic root"""
    for i, k in enumerate(params["k_is"]):
        assert O.gpow(7, i) == k
    g = O.gpow(7, (P - 1) >> 32)
    assert g == 1753635133440165772
    assert O.gpow(g, 1 << 31) == P - 1 and O.gpow(g, 1 << 32) == 1
    for k in (1, 3, 12, 15, 20, 23):
        w = O.root_of_unity(k)
        assert O.gpow(w, 1 << k) == 1 and O.gpow(w, 1 << (k - 1)) == P - 1


def test_k5_params_match_stored_proofs(params, proofs):
    """zk_signature2/mod.rs:33-57 vs the shapes of all ten stored proofs"""
    nq = params["num_query_rounds"]
    widths = params["num_leaves_per_oracle"]
    ncap = 1 << params["cap_height"]
    lde_bits = params["degree_bits"] + params["rate_bits"]
    for p in proofs:
        assert p["wires_cap"].shape == (ncap, 4) and p["zs_pp_cap"].shape == (ncap, 4)
        assert len(p["query_rounds"]) == nq
        assert len(p["commit_phase_merkle_caps"]) == len(params["reduction_arity_bits"])
        r0 = p["query_rounds"][0]
        assert [len(l) for l, _ in r0["initial"]] == widths
        assert all(s.shape[0] == lde_bits - params["cap_height"] for _, s in r0["initial"])
        bits = lde_bits
        for (ev, sib), ab in zip(r0["steps"], params["reduction_arity_bits"]):
            bits -= ab
            assert ev.shape == (1 << ab, 2) and sib.shape[0] == bits - params["cap_height"]
        assert p["final_poly"].shape[0] == 1 << (params["degree_bits"] - sum(params["reduction_arity_bits"]))
        assert p["openings"]["wires"].shape[0] == params["num_wires"]
        assert p["openings"]["plonk_sigmas"].shape[0] == params["num_routed_wires"]
        assert p["openings"]["constants"].shape[0] == params["num_constants"]
        assert p["openings"]["quotient_polys"].shape[0] == params["num_quotient_polys"]
        assert p["openings"]["partial_products"].shape[0] == params["total_partial_products"]
        for k, v in p["openings"].items():
            assert (v < P).all(), k  # serialised felts are canonical


def test_k3_merkle_paths_of_stored_proofs(params, proofs):
    """Every query round of every stored proof: the 135/20/16-wide initial leaves and the 32-felt FRI
    layer leaves hash (hash_or_noop sponge) and climb (two_to_one, bit=1 => H(sib||cur)) to
    cap[index >> n_siblings] of the caps carried in the same proof."""
    caps_of = lambda p: [None, p["wires_cap"], p["zs_pp_cap"], p["quotient_cap"]]
    checked = 0
    for p in proofs:
        caps = caps_of(p)
        for rnd in p["query_rounds"]:
            leaf, sib = rnd["initial"][3]
            x_index = O.merkle_find_index(leaf, sib, caps[3])
            assert x_index >= 0, "no unique query index"
            for t in (1, 2, 3):
                leaf, sib = rnd["initial"][t]
                assert O.merkle_verify(leaf, x_index, sib, caps[t])
                # and a wrong index must fail
                assert not O.merkle_verify(leaf, x_index ^ 1, sib, caps[t])
            xi = x_index
            for (ev, sib), cap, ab in zip(rnd["steps"], p["commit_phase_merkle_caps"], params["reduction_arity_bits"]):
                xi >>= ab
                assert O.merkle_verify(ev.reshape(-1), xi, sib, cap)
            checked += 1
    assert checked == 10 * params["num_query_rounds"]


def test_field_mul_against_slow_mod():
    rng = np.random.default_rng(1)
    L = O.lib()
    edge = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32, 2**32 - 1, 2**63]
    vals = edge + [int(x) for x in rng.integers(0, 2**64, 200, dtype=np.uint64)]
    for a in vals[:40]:
        for b in vals:
            assert L.gl_mul(a, b) == (a * b) % P == L.gl_mul_slow(a, b)
            assert L.gl_add(a, b) == (a + b) % P
            assert L.gl_sub(a, b) == (a - b) % P
    for a in vals:
        if a % P:
            assert L.gl_mul(a, L.gl_inv(a)) == 1
