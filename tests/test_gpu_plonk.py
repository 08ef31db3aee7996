"""GPU parity of the PLONK stages between the commitments through the C ABI:
p2b_zs_partial_products_commit and p2b_quotient_commit against the oracle (oracle/plonk.c) — which
tests/test_plonk_oracle.py pins to the independent verifier-identity restatement — on synthetic circuits over
the closed gate set, bit-exact: Z / partial-product values, their Merkle cap, the quotient chunk coefficients
and the quotient commitment's cap.  The largest case also re-checks the verifier identity on the GPU's own
output."""
import random

import numpy as np
import pytest

import p2oracle as O
import plonk_ref as R
from test_plonk_oracle import (ALL_GATES, COSET_GATES, COSET_GROUPS, EXT_GATES, EXT_GROUPS, MORE_GATES, MORE_GROUPS,
                               check_verifier_identity)

pytestmark = pytest.mark.gpu
P = R.P
FULL_GROUPS = [(0, 4), (4, 5), (5, 8), (8, 10)]


@pytest.fixture(scope="module")
def m():
    import city_rollup_b200 as mod

    mod.load()
    return mod


@pytest.fixture(scope="module")
def ctx(m):
    c = m.Context(0)
    yield c
    c.close()


def run_gpu(ctx, m, circ, betas, gammas, alphas, rate_bits, cap_height):
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), rate_bits, False, cap_height, keep_values=True)
    wi = m.PolynomialBatch.from_values(ctx, circ.wire_values(), rate_bits, False, cap_height, keep_values=True)
    zs = m.all_wires_permutation_partial_products(ctx, cd, cs, wi, betas, gammas, rate_bits, cap_height)
    qt = m.compute_quotient_polys(ctx, cd, cs, circ.pi_hash, wi, zs, betas, gammas, alphas, rate_bits, cap_height)
    return cd, cs, wi, zs, qt


@pytest.mark.parametrize("degree_bits,gates,groups,seed,cap_height", [
    (3, ALL_GATES[:4], [(0, 4)], 21, 0),
    (5, ALL_GATES[:5], [(0, 4), (4, 5)], 22, 2),
    (6, ALL_GATES, FULL_GROUPS, 23, 4),
    (9, ALL_GATES, FULL_GROUPS, 24, 4),
    (7, MORE_GATES, MORE_GROUPS, 26, 3),
    (7, EXT_GATES, EXT_GROUPS, 27, 4),
    (6, COSET_GATES, COSET_GROUPS, 28, 2),
    (12, ALL_GATES, FULL_GROUPS, 25, 4),
])
def test_plonk_stages_match_oracle(ctx, m, degree_bits, gates, groups, seed, cap_height):
    check_plonk_stages(ctx, m, R.SyntheticCircuit(degree_bits, gates, groups, seed), 3, cap_height, seed)


@pytest.mark.parametrize("degree_bits,qdf,rate_bits,gates,groups,seed", [
    # quotient_degree_factor 4: only the degree-2 gates can be evaluated on a sub-coset (2n of 4n points), the
    # degree-3 ArithmeticGate is evaluated at every point
    (6, 4, 3, [(R.GATE_CONSTANT, 2, 0), (R.GATE_BASE_SUM, 10, 0), (R.GATE_ARITHMETIC, 20, 0)], [(0, 2), (2, 3)], 41),
    # quotient_degree_factor 2: no sub-coset is smaller than the quotient coset
    (5, 2, 3, [(R.GATE_CONSTANT, 2, 0), (R.GATE_PUBLIC_INPUT, 0, 0)], [(0, 1), (1, 2)], 42),
    # LDE rate 16 with quotient_degree_factor 8: the quotient coset is itself a sub-coset of the batches' LDE
    (6, 8, 4, ALL_GATES, FULL_GROUPS, 43),
])
def test_plonk_stages_other_degree_factors(ctx, m, degree_bits, qdf, rate_bits, gates, groups, seed):
    circ = R.SyntheticCircuit(degree_bits, gates, groups, seed, quotient_degree_factor=qdf)
    check_plonk_stages(ctx, m, circ, rate_bits, 2, seed)


def check_plonk_stages(ctx, m, circ, rate_bits, cap_height, seed):
    degree_bits = circ.degree_bits
    rng = random.Random(seed + 1)
    nch = circ.num_challenges
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(nch)] for _ in range(3))
    d = circ.desc()
    cd, cs, wi, zs, qt = run_gpu(ctx, m, circ, betas, gammas, alphas, rate_bits, cap_height)
    # ---- a6: Z and partial products
    ref_zs = O.partial_products_and_zs(d, np.array(circ.wires, dtype=np.uint64), np.array(circ.sigmas, dtype=np.uint64),
                                       betas, gammas)
    assert zs.n_cols == ref_zs.shape[0] == nch * (1 + circ.num_pp)
    for c in range(zs.n_cols):
        assert (zs.values(c) == ref_zs[c]).all(), f"Z / partial product column {c}"
    o_cs = O.batch_from_values(circ.constants_sigmas_values(), rate_bits, cap_height, want_digests=False)
    o_wi = O.batch_from_values(circ.wire_values(), rate_bits, cap_height, want_digests=False)
    o_zs = O.batch_from_values(list(ref_zs), rate_bits, cap_height, want_digests=False)
    assert (zs.cap == o_zs["cap"]).all()
    # ---- a7 / a8: quotient chunks and their commitment
    ref_chunks = O.compute_quotient_polys(d, rate_bits, o_cs["leaves"], o_wi["leaves"], o_zs["leaves"], circ.pi_hash,
                                          betas, gammas, alphas)
    assert qt.n_cols == ref_chunks.shape[0] == nch * circ.qdf
    got = np.stack([qt.coeffs(c) for c in range(qt.n_cols)])
    assert (got == ref_chunks).all(), "quotient chunk coefficients"
    o_q = O.batch_from_coeffs(list(ref_chunks), rate_bits, cap_height, want_leaves=False, want_digests=False)
    assert (qt.cap == o_q["cap"]).all()
    if degree_bits == 6:  # the GPU's own output satisfies the verifier identity (independent restatement)
        pr = dict(betas=betas, gammas=gammas, alphas=alphas, chunks=got,
                  cs=dict(coeffs=np.stack([cs.coeffs(c) for c in range(cs.n_cols)])),
                  wires=dict(coeffs=np.stack([wi.coeffs(c) for c in range(wi.n_cols)])),
                  zs=dict(coeffs=np.stack([zs.coeffs(c) for c in range(zs.n_cols)])))
        check_verifier_identity(circ, pr, seed + 5)
    for h in (qt, zs, wi, cs, cd):
        h.free()


def test_plonk_argument_checks(ctx, m):
    circ = R.SyntheticCircuit(3, ALL_GATES[:4], [(0, 4)], 31)
    d = circ.desc()
    bad = dict(d, num_partial_products=3)
    with pytest.raises(m.P2BError):
        m.CircuitData(ctx, bad)
    bad = dict(d, gates=[dict(d["gates"][0], kind=99)] + d["gates"][1:])
    with pytest.raises(m.P2BError):
        m.CircuitData(ctx, bad)
    cd = m.CircuitData(ctx, d)
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 0)  # values not kept
    wi = m.PolynomialBatch.from_values(ctx, circ.wire_values(), 3, False, 0, keep_values=True)
    with pytest.raises(m.P2BError) as e:
        m.all_wires_permutation_partial_products(ctx, cd, cs, wi, [1, 2], [3, 4], 3, 0)
    assert "KEEP_VALUES" in str(e.value)
    with pytest.raises(m.P2BError):  # wrong width
        m.all_wires_permutation_partial_products(ctx, cd, wi, wi, [1, 2], [3, 4], 3, 0)
    for h in (wi, cs, cd):
        h.free()
