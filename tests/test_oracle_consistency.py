"""CPU-only self-consistency of the oracle's unpinned parts (NTT/LDE/FRI/challenger): the reference holds
no golden vector for these, so they are checked against independent definitions."""
import numpy as np

import p2oracle as O
from util import P, bitrev, rand_felts


def horner(coeffs, x):
    acc = 0
    for c in reversed([int(c) for c in coeffs]):
        acc = (acc * x + c) % P
    return acc


def test_fft_matches_direct_evaluation():
    for log_n in (0, 1, 2, 5, 8):
        n = 1 << log_n
        c = rand_felts(100 + log_n, n)
        v = O.fft(c)
        w = O.root_of_unity(log_n)
        for r in {0, 1 % n, n // 2, n - 1, (3 * n) // 7}:
            assert int(v[r]) == horner(c, pow(w, r, P))
        assert (O.ifft(v) == c).all()


def test_coset_fft_and_lde_rows():
    log_n, rate = 6, 3
    n, N = 1 << log_n, 1 << (log_n + rate)
    cols = [rand_felts(7 + i, n) for i in range(5)]
    out = O.batch_from_coeffs(cols, rate, 2)
    w = O.root_of_unity(log_n + rate)
    for j in (0, 1, 17, N - 1):
        x = 7 * pow(w, bitrev(j, log_n + rate), P) % P
        for c in range(5):
            assert int(out["leaves"][j, c]) == horner(cols[c], x)
    vals = [O.fft(c) for c in cols]
    out2 = O.batch_from_values(vals, rate, 2)
    assert (out2["coeffs"] == np.array(cols)).all()
    assert (out2["leaves"] == out["leaves"]).all() and (out2["cap"] == out["cap"]).all()


def test_merkle_prove_verify_and_layout():
    for (n, w, ch) in ((16, 7, 2), (8, 3, 0), (32, 20, 5), (4, 9, 1), (64, 135, 4)):
        leaves = rand_felts(n * 1000 + w, (n, w))
        dg, cap = O.merkle_tree_new(leaves, ch)
        assert dg.shape[0] == 2 * (n - (1 << ch))
        for i in range(n):
            sib = O.merkle_prove(dg, n, ch, i)
            assert O.merkle_verify(leaves[i], i, sib, cap)
        # cap_height == log2(n): cap = leaf digests
        if (1 << ch) == n:
            for i in range(n):
                assert (cap[i] == O.hash_or_noop(leaves[i])).all()


def test_challenger_is_a_duplex_sponge():
    ch = O.Challenger()
    ch.observe([1, 2, 3])
    a = ch.get()
    s = O.permute([1, 2, 3] + [0] * 9)
    assert a == int(s[7])  # pops from the end of the squeezed rate
    assert ch.get() == int(s[6])
    ch.observe(list(range(10, 19)))  # 9 elements: one automatic duplex at 8, one pending
    s2 = s.copy()
    s2[:8] = np.arange(10, 18, dtype=np.uint64)
    s2 = O.permute(s2)
    s3 = s2.copy()
    s3[0] = 18
    s3 = O.permute(s3)
    assert ch.get() == int(s3[7])


def test_fri_fold_matches_verifier_interpolation():
    """prover-side fold of coefficients == verifier-side compute_evaluation on the committed leaves"""
    log_n, rate = 7, 3
    n = 1 << log_n
    N = n << rate
    coeffs = np.zeros((N, 2), np.uint64)
    coeffs[:n] = rand_felts(5, (n, 2))
    values = O.ext_coset_fft(coeffs, 7)
    ch = O.Challenger()
    ch.observe([42])
    arity_bits = [4, 3]
    out = O.fri_committed_trees(coeffs, values, arity_bits, ch, rate, 1)
    log_N = log_n + rate
    w = O.root_of_unity(log_N)
    for x_index in (0, 5, 321, N - 1):
        xi = x_index
        x = 7 * pow(w, bitrev(xi, log_N), P) % P
        leaf = out["leaves"][0][xi >> 4].reshape(16, 2)
        ev = O.fri_compute_evaluation(x, xi & 15, 4, leaf, out["betas"][0])
        xi >>= 4
        x = pow(x, 16, P)
        leaf1 = out["leaves"][1][xi >> 3].reshape(8, 2)
        assert [int(v) for v in leaf1[xi & 7]] == ev
        ev2 = O.fri_compute_evaluation(x, xi & 7, 3, leaf1, out["betas"][1])
        x = pow(x, 8, P)
        # final poly evaluated at x (extension coefficients, base point)
        acc = [0, 0]
        for c in reversed(out["final_poly"].tolist()):
            acc = [(acc[0] * x + c[0]) % P, (acc[1] * x + c[1]) % P]
        assert acc == ev2
    assert out["final_poly"].shape[0] == (N >> 7) >> rate


def test_pow_is_minimal_and_valid():
    ch = O.Challenger()
    ch.observe([9, 8, 7])
    base = ch.clone()
    w = O.fri_proof_of_work(ch, 10)
    assert O.fri_pow_check(base, w, 10) == 1
    assert all(O.fri_pow_check(base, v, 10) == 0 for v in range(w))


def test_poseidon_fast_partial_round_form_equals_permutation():
    """the fast-partial-round schedule (tables derived in tools/gen_poseidon_fast_tables.py; the form PoseidonGate
    constrains) computes the K1-pinned permutation"""
    L = O.lib()
    rng = np.random.default_rng(11)
    for t in range(300):
        s = rng.integers(0, 2**64, 12, dtype=np.uint64) if t > 1 else np.full(12, 2**64 - 1 if t else 0, dtype=np.uint64)
        a, b = s.copy(), s.copy()
        L.poseidon_permute(O._p(a))
        L.poseidon_permute_fast(O._p(b))
        assert (a == b).all()


def test_linearised_partial_round_tables_are_the_generated_ones(tmp_path):
    """city_rollup_b200/csrc/poseidon_partial_lin.inc (the one-warp permutation's 21 linearised partial rounds) is
    exactly what tools/gen_poseidon_partial_linear.py derives from the K1-pinned round constants and the MDS matrix
    (the generator checks its forms against the naive partial rounds before writing)."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_pl", os.path.join(root, "tools", "gen_poseidon_partial_linear.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    xs, outs = g.derive()
    assert len(xs) == 21 and len(outs) == 12
    import random

    rng = random.Random(5)
    s = [rng.randrange(g.P) for _ in range(12)]
    assert g.partial_rounds_linear(s, xs, outs) == g.partial_rounds_naive(s)
    committed = open(os.path.join(root, "city_rollup_b200", "csrc", "poseidon_partial_lin.inc")).read()
    forms = [xs[l + 1] for l in range(20)] + outs
    # spot-check the committed table against the derivation: constants, first and last S-box coefficient rows
    assert "#define PL_X0_CONST 0x%016xull" % xs[0][-1] in committed
    for l in (0, 19, 20, 31):
        assert "0x%016xull" % forms[l][-1] in committed
        assert "0x%016xull" % forms[l][12] in committed
    # x_{l+1} must not depend on S-box outputs it cannot have seen
    for l in range(20):
        assert all(forms[l][12 + k] == 0 for k in range(l + 1, 21))


def test_v6_poseidon_schedule_model_and_tables():
    """The per-thread permutation schedule of poseidon.cuh (unreduced last S-box product as limbs, chained MDS layers of
    a partial-round pair, lazy folds) restated operation by operation in exact integer arithmetic
    (tools/gen_poseidon_v6_tables.py) equals the plain permutation — the oracle's, pinned by K1/K2 — on random and
    extreme states, every FP64 intermediate stays exact with the limb values at the corners of their ranges, and the
    committed chain initialisers are the generated ones."""
    import importlib.util
    import os
    import random

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_v6", os.path.join(root, "tools", "gen_poseidon_v6_tables.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    rng = random.Random(7)
    states = [[0] * 12, [2**64 - 1] * 12, [g.P - 1] * 12] + [[rng.getrandbits(64) for _ in range(12)] for _ in range(40)]
    for st in states:
        got = [x % g.P for x in g.permute_v6(st)]
        assert got == g.permute_ref(st)
        assert got == [int(x) for x in O.permute([x % g.P for x in st])]
        # the Q schedule (a partial-round pair as one application of M^2): what permute_nc runs
        assert [x % g.P for x in g.permute_q(st)] == got
    # ... and with the S-box input of the second round of every pair replaced (the PoseidonGate evaluator's hook): the
    # pair must equal two plain rounds with the same replacement
    for _ in range(20):
        r = rng.choice(range(4, 26, 2))
        s0, vals, repl = rng.getrandbits(64), [rng.getrandbits(64) for _ in range(11)], rng.randrange(g.P)
        lz = [None] + [((v & 0xFFFFFFFF) + g.OL, v >> 32) for v in vals]
        st = [s0 % g.P] + [v % g.P for v in vals]
        mds = lambda s: [(sum(g.CIRC[(k - i) % 12] * s[k] for k in range(12)) + (8 * s[0] if i == 0 else 0)) % g.P for i in range(12)]
        st[0] = pow(st[0], 7, g.P)
        st = [(x + c) % g.P for x, c in zip(mds(st), g.RC[12 * (r + 1):12 * (r + 2)])]
        computed_x2 = st[0]
        st[0] = pow(repl, 7, g.P)
        st = [(x + c) % g.P for x, c in zip(mds(st), g.RC[12 * (r + 2):12 * (r + 3)])]
        seen = []
        s0n, lzn = g.pair_q(s0, lz, r, mid=lambda x: (seen.append(x), repl)[1])
        assert seen[0] % g.P == computed_x2
        assert [s0n % g.P] + [(lzn[i][0] - g.OL + (lzn[i][1] << 32)) % g.P for i in range(1, 12)] == st
    # committed table == generated table
    committed = open(os.path.join(root, "city_rollup_b200", "csrc", "poseidon_rc_v6.inc")).read()
    words = [int(x, 16) for x in __import__("re").findall(r"0x([0-9a-f]{16})ull", committed)]
    want = []
    for r in range(30):
        for limb in range(2):
            for rr in range(6):
                a, b = g.INITS[r][limb][rr]
                want += [g.bits(a), g.bits(b)]
    assert words == want
    committed = open(os.path.join(root, "city_rollup_b200", "csrc", "poseidon_rc_q.inc")).read()
    words = [int(x, 16) for x in __import__("re").findall(r"0x([0-9a-f]{16})ull", committed)]
    want = []
    for r in range(4, 26, 2):
        x2, oi = g.Q_INITS[r]
        for limb in range(2):
            want += [g.bits(v) for v in [x2[limb][0], x2[limb][1]] + [oi[limb][rr][0] for rr in range(6)] + [oi[limb][rr][1] for rr in range(6)]]
    assert words == want
    # FP64 bounds at the corners of the limb ranges (asserts inside the model)
    real_sbox, real_lazy = g.sbox_limbs, g.lazy_fold
    try:
        g.sbox_limbs = lambda x: (rng.choice([0, 2**33 + 2**32 - 1]), rng.choice([0, 2**33 - 2]))

        def corner_lazy(ya, yb):
            real_lazy(ya, yb)
            return rng.choice([0, 2**32 + 2**18 - 1]), rng.choice([0, 2**32 + 2**19])

        g.lazy_fold = corner_lazy
        for _ in range(200):
            g.permute_v6([0] * 12)
            g.permute_q([0] * 12)
    finally:
        g.sbox_limbs, g.lazy_fold = real_sbox, real_lazy
