"""Pin the CPU oracle to every known answer the reference tree holds for the hot path
(SURVEY.md §8(c) K1..K6).  CPU-only."""
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


def test_k6_sighash_whitelist_root(golden_dir):
    """city_rollup_common/src/config/sighash_wrapper_config.rs:14-23: SIGHASH_WHITELIST_TREE_ROOT is the root of a height-16
    zero-hash Merkle tree over the 1 875 circuit fingerprints of :24-1900 (city_store/src/store/sighash/mod.rs:44-75).  A
    4-element leaf is its own digest (hash_or_noop) and an empty subtree is the K1 zero hash, so the root is the one-element
    cap of MerkleTree::new over the 2^16 leaves: 65 535 two_to_one compressions over real data, against a root the reference
    holds as a constant."""
    from util import sighash_whitelist_leaves

    wl = json.load(open(os.path.join(golden_dir, "sighash_whitelist.json")))
    leaves = sighash_whitelist_leaves(wl)
    digests, cap = O.merkle_tree_new(leaves, 0)
    assert cap.shape == (1, 4) and cap[0].tolist() == wl["root"]
    # a whitelist inclusion proof as the wrapper circuit checks it (sighash_wrapper.rs:73-80): 16 siblings up to the root
    j = 1234
    siblings = O.merkle_prove(digests, len(leaves), 0, j)
    assert len(siblings) == wl["tree_height"] and O.merkle_verify(leaves[j], j, siblings, cap)


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


# --------------------------------------------------------------------------------------------------
# K3, FRI part: the fold convention (point order inside a layer leaf, coset shift, arity fold) pinned on
# the stored proofs.  The FRI betas depend on the transcript (circuit digest not in the dump), so they are
# SOLVED: in every query round the degree-15 interpolant P_q of the 16 opened layer-0 evaluations satisfies
# P_q(beta_0) = (opened layer-1 evaluation); two rounds give beta_0 as the common root (polynomial gcd over
# GF(p^2)), and the remaining 26 rounds — and the oracle's fri_compute_evaluation — must agree with it.
# The same is done for beta_1 against final_poly.
def _emul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _esub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def _einv(a):
    n = pow((a[0] * a[0] - 7 * a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def _interp_coeffs(xs, ys):
    """coefficients (ext) of the polynomial through (xs[i] base, ys[i] ext)"""
    n = len(xs)
    master = [1]
    for x in xs:  # master(X) *= (X - x)
        master = [(-x * master[0]) % P] + [(master[i - 1] - x * master[i]) % P for i in range(1, len(master))] + [master[-1]]
    out = [(0, 0)] * n
    for i in range(n):
        q = [0] * n  # master / (X - xs[i]) by synthetic division
        q[n - 1] = master[n]
        for k in range(n - 1, 0, -1):
            q[k - 1] = (master[k] + xs[i] * q[k]) % P
        d = 1
        for j in range(n):
            if j != i:
                d = d * (xs[i] - xs[j]) % P
        s = pow(d, P - 2, P)
        out = [((o[0] + ys[i][0] * q[k] % P * s) % P, (o[1] + ys[i][1] * q[k] % P * s) % P) for k, o in enumerate(out)]
    return out


def _poly_trim(a):
    while a and a[-1] == (0, 0):
        a = a[:-1]
    return a


def _poly_mod(a, b):
    a = _poly_trim(list(a))
    b = _poly_trim(list(b))
    inv = _einv(b[-1])
    while len(a) >= len(b):
        f = _emul(a[-1], inv)
        sh = len(a) - len(b)
        for i, c in enumerate(b):
            a[sh + i] = _esub(a[sh + i], _emul(f, c))
        a = _poly_trim(a)
    return a


def _poly_gcd(a, b):
    while _poly_trim(list(b)):
        a, b = b, _poly_mod(a, b)
    return _poly_trim(list(a))


def _coset_points(x, within, arity_bits):
    arity = 1 << arity_bits
    g = O.root_of_unity(arity_bits)
    rev = int(format(within, f"0{arity_bits}b")[::-1], 2)
    start = x * pow(g, arity - rev, P) % P
    return [start * pow(g, i, P) % P for i in range(arity)]


def _bitrev_list(v, bits):
    return [v[int(format(i, f"0{bits}b")[::-1], 2)] for i in range(len(v))]


def test_k3_fri_fold_convention_on_stored_proofs(params, proofs):
    log_N = params["degree_bits"] + params["rate_bits"]
    w = O.root_of_unity(log_N)
    for p in proofs[:3]:
        rounds = p["query_rounds"]
        data = []
        for rnd in rounds:
            x_index = O.merkle_find_index(rnd["initial"][3][0], rnd["initial"][3][1], p["quotient_cap"])
            x0 = 7 * pow(w, int(format(x_index, f"0{log_N}b")[::-1], 2), P) % P
            data.append((x_index, x0))

        def layer_poly(rnd, x, within, layer):
            ev = [tuple(int(c) for c in e) for e in rnd["steps"][layer][0]]
            return _interp_coeffs(_coset_points(x, within, 4), _bitrev_list(ev, 4))

        # ---- beta_0 from rounds 0 and 1
        polys = []
        for q in (0, 1):
            xi, x0 = data[q]
            pol = layer_poly(rounds[q], x0, xi & 15, 0)
            target = tuple(int(c) for c in rounds[q]["steps"][1][0][(xi >> 4) & 15])
            pol[0] = _esub(pol[0], target)
            polys.append(pol)
        g = _poly_gcd(polys[0], polys[1])
        assert len(g) == 2, "expected a unique common root"
        beta0 = _emul(((-g[0][0]) % P, (-g[0][1]) % P), _einv(g[1]))
        # ---- beta_1 from final_poly
        fin = [tuple(int(c) for c in e) for e in p["final_poly"]]

        def final_eval(x):
            acc = (0, 0)
            for c in reversed(fin):
                acc = ((acc[0] * x + c[0]) % P, (acc[1] * x + c[1]) % P)
            return acc

        polys = []
        for q in (0, 1):
            xi, x0 = data[q]
            x1 = pow(x0, 16, P)
            pol = layer_poly(rounds[q], x1, (xi >> 4) & 15, 1)
            pol[0] = _esub(pol[0], final_eval(pow(x1, 16, P)))
            polys.append(pol)
        g = _poly_gcd(polys[0], polys[1])
        assert len(g) == 2
        beta1 = _emul(((-g[0][0]) % P, (-g[0][1]) % P), _einv(g[1]))
        # ---- every round, through the oracle's verifier-side fold
        for rnd, (xi, x0) in zip(rounds, data):
            e0 = O.fri_compute_evaluation(x0, xi & 15, 4, rnd["steps"][0][0], list(beta0))
            assert e0 == [int(c) for c in rnd["steps"][1][0][(xi >> 4) & 15]]
            x1 = pow(x0, 16, P)
            e1 = O.fri_compute_evaluation(x1, (xi >> 4) & 15, 4, rnd["steps"][1][0], list(beta1))
            assert tuple(e1) == final_eval(pow(x1, 16, P))


# --------------------------------------------------------------------------------------------------
# K3, opening part: the verifier's fri_combine_initial on the stored proofs.  It depends on two transcript
# values that cannot be recomputed (the circuit digest is not in the dump): the FRI batching challenge alpha and
# the opening point zeta.  Both are SOLVED from the proofs themselves: with U(a) = a^2 sum_i a^i (p_i(x) - y_i)
# over the 256 polynomials opened at zeta (oracle order constants|sigmas, wires, Zs|partial products, quotient)
# and V(a) = sum_j a^j (Z_j(x) - z_next_j), every query round gives
#       E (x - zeta)(x - g zeta) = U (x - g zeta) + V (x - zeta),       E = the opened layer-0 FRI evaluation,
# a quadratic in zeta whose coefficients are polynomials in alpha.  Three rounds make a 3x3 determinant that
# must vanish at the true alpha; the gcd of two such determinants (degree ~500 over GF(p^2)) is linear and gives
# alpha, two rounds then give zeta, and ALL 28 rounds must satisfy the identity.  This pins the opening order,
# the two opening batches, the alpha bookkeeping (reduce / shift), the point convention x = 7 w^bitrev(index)
# and g = primitive n-th root — i.e. exactly what the prover's prove_openings has to produce.
def _padd(a, b):
    n = max(len(a), len(b))
    a = a + [(0, 0)] * (n - len(a))
    b = b + [(0, 0)] * (n - len(b))
    return [((x[0] + y[0]) % P, (x[1] + y[1]) % P) for x, y in zip(a, b)]


def _pscale(a, s):
    return [_emul(x, s) for x in a]


def _pmul(a, b):
    out = [[0, 0] for _ in range(len(a) + len(b) - 1)]
    for i, x in enumerate(a):
        if x == (0, 0):
            continue
        x0, x1 = x
        for j, y in enumerate(b):
            o = out[i + j]
            o[0] += x0 * y[0] + 7 * x1 * y[1]
            o[1] += x0 * y[1] + x1 * y[0]
    return [(o[0] % P, o[1] % P) for o in out]


def _peval(a, x):
    acc = (0, 0)
    for c in reversed(a):
        acc = _emul(acc, x)
        acc = ((acc[0] + c[0]) % P, (acc[1] + c[1]) % P)
    return acc


@pytest.mark.parametrize("which", [0, 3])  # a circuit-64 (zk signature) and a circuit-65 (secp256k1) proof
def test_k3_fri_combine_initial_on_stored_proof(params, proofs, which):
    p = proofs[which]
    log_N = params["degree_bits"] + params["rate_bits"]
    w = O.root_of_unity(log_N)
    g = O.root_of_unity(params["degree_bits"])
    op = p["openings"]
    ys = [tuple(int(c) for c in e) for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products",
                                             "quotient_polys") for e in op[k]]
    ys_next = [tuple(int(c) for c in e) for e in op["plonk_zs_next"]]
    nch = len(ys_next)
    assert len(ys) == sum(params["num_leaves_per_oracle"])
    # the constants|sigmas leaves (oracle 0) verify against no cap we hold, but they are needed: take them as given
    rows = []
    for rnd in p["query_rounds"]:
        x_index = O.merkle_find_index(rnd["initial"][3][0], rnd["initial"][3][1], p["quotient_cap"])
        x = 7 * pow(w, int(format(x_index, f"0{log_N}b")[::-1], 2), P) % P
        evals = [int(v) for leaf, _ in rnd["initial"] for v in leaf]
        zs_leaf = [int(v) for v in rnd["initial"][2][0][:nch]]
        E = tuple(int(c) for c in rnd["steps"][0][0][x_index & 15])
        U = [(0, 0), (0, 0)] + [((e - y[0]) % P, (-y[1]) % P) for e, y in zip(evals, ys)]
        V = [((e - y[0]) % P, (-y[1]) % P) for e, y in zip(zs_leaf, ys_next)]
        a = _emul(E, (g, 0))
        b = _padd(_padd(_pscale(U, (g, 0)), V), [_emul(E, ((-x * (1 + g)) % P, 0))])
        c = _padd(_pscale(_padd(U, V), ((-x) % P, 0)), [_emul(E, (x * x % P, 0))])
        rows.append((a, b, c, x, E, U, V))

    def det(i, j, k):
        (a1, b1, c1), (a2, b2, c2), (a3, b3, c3) = (rows[t][:3] for t in (i, j, k))
        neg = lambda poly: [((-u[0]) % P, (-u[1]) % P) for u in poly]
        m1 = _padd(_pmul(b2, c3), neg(_pmul(b3, c2)))
        m2 = _padd(_pmul(b1, c3), neg(_pmul(b3, c1)))
        m3 = _padd(_pmul(b1, c2), neg(_pmul(b2, c1)))
        return _padd(_padd(_pscale(m1, a1), neg(_pscale(m2, a2))), _pscale(m3, a3))

    gg = _poly_gcd(det(0, 1, 2), det(0, 1, 3))
    assert len(gg) == 2, f"expected a unique common root, gcd has degree {len(gg) - 1}"
    alpha = _emul(((-gg[0][0]) % P, (-gg[0][1]) % P), _einv(gg[1]))
    # zeta from rounds 0 and 1: (a1 b2 - a2 b1) zeta + (a1 c2 - a2 c1) = 0
    a1, b1, c1 = rows[0][0], _peval(rows[0][1], alpha), _peval(rows[0][2], alpha)
    a2, b2, c2 = rows[1][0], _peval(rows[1][1], alpha), _peval(rows[1][2], alpha)
    lin = _esub(_emul(a1, b2), _emul(a2, b1))
    con = _esub(_emul(a1, c2), _emul(a2, c1))
    zeta = _emul(((-con[0]) % P, (-con[1]) % P), _einv(lin))
    # zeta is outside H (plonky2 aborts otherwise) and every one of the 28 rounds agrees
    zn = (1, 0)
    zp = zeta
    for _ in range(params["degree_bits"]):
        zp = _emul(zp, zp)
    assert zp != (1, 0)
    for a, b, c, x, E, U, V in rows:
        u, v = _peval(U, alpha), _peval(V, alpha)
        s = ((x - zeta[0]) % P, (-zeta[1]) % P)
        gz = _emul(zeta, (g, 0))
        t = ((x - gz[0]) % P, (-gz[1]) % P)
        total = _emul(u, _einv(s))
        nxt = _emul(v, _einv(t))
        assert ((total[0] + nxt[0]) % P, (total[1] + nxt[1]) % P) == E
