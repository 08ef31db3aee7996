"""GPU parity tests: the CUDA path, driven through the C ABI (libp2b.so), against the CPU oracle on the
same seeded inputs, against the reference's golden fixtures, and — at BASELINE.json's full sizes —
through size-independent properties.  Bit-exact everywhere (integer arithmetic mod p)."""
import json
import os

import numpy as np
import pytest

import p2oracle as O
from proof_parser import parse_proof
from util import P, bitrev, rand_felts, splitmix64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import city_rollup_b200 as m

    c = m.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def m():
    import city_rollup_b200 as mod

    return mod


# ----------------------------------------------------------------------------- Poseidon
def test_poseidon_permutation_random_and_noncanonical(ctx):
    states = rand_felts(11, (300, 12), canonical=False)
    states[0] = 0
    states[1] = np.uint64(P - 1)
    states[2] = np.uint64(2**64 - 1)
    states[3] = np.uint64(P)
    got = ctx.poseidon_permute(states)
    for i in range(states.shape[0]):
        assert (got[i] == O.permute(states[i])).all(), i
    assert (got < np.uint64(P)).all()


def test_poseidon_permutation_carry_patterns(ctx):
    """states drawn from the words that stress the carry / borrow paths of the limb-form schedule (all-ones and
    all-zero 32-bit halves, p - 1, p, 2^64 - 1), plus a larger random batch; bit-exact against the oracle"""
    rng = np.random.default_rng(23)
    words = np.array([0, 1, 2**32 - 1, 2**32, 2**32 + 1, 2**64 - 2**32, P - 1, P, P + 1, 2**64 - 1, 2**63, 2**33 - 1],
                     dtype=np.uint64)
    states = np.concatenate([words[rng.integers(0, len(words), (2048, 12))],
                             rand_felts(29, (4096, 12), canonical=False)])
    got = ctx.poseidon_permute(states)
    for i in range(states.shape[0]):
        assert (got[i] == O.permute(states[i])).all(), i


def test_k1_k2_zero_hash_chains_on_gpu(ctx, golden_dir):
    """city_crypto/src/hash/cached_zero_hashes.rs:10-1036 and :1039-2066, computed by the CUDA kernels"""
    z = json.load(open(os.path.join(golden_dir, "zero_hashes.json")))
    cur = np.zeros((1, 4), np.uint64)
    for i in range(1, 128):
        cur = ctx.two_to_one(cur, cur)
        assert cur[0].tolist() == z["zero"][i], i
    cur = ctx.hash_no_pad([0] * 8 + [1]).reshape(1, 4)
    assert cur[0].tolist() == z["marked"][1]
    for i in range(2, 128):
        cur = ctx.two_to_one(cur, cur)
        assert cur[0].tolist() == z["marked"][i], i


def test_k6_sighash_whitelist_root_on_gpu(ctx, m, golden_dir):
    """K6 on the GPU, against the constant the reference holds (not against the oracle): MerkleTree::new over the 2^16
    leaves of the sighash-circuit whitelist (1 875 real fingerprints + zero hashes), cap_height 0 ->
    SIGHASH_WHITELIST_TREE_ROOT (city_rollup_common/src/config/sighash_wrapper_config.rs:14-23); the inclusion proof of one
    fingerprint climbs to it."""
    from util import sighash_whitelist_leaves

    wl = json.load(open(os.path.join(golden_dir, "sighash_whitelist.json")))
    leaves = sighash_whitelist_leaves(wl)
    t = m.MerkleTree.new(ctx, leaves, 0)
    assert t.cap.shape == (1, 4) and t.cap[0].tolist() == wl["root"]
    j = 777
    assert O.merkle_verify(leaves[j], j, t.prove(j), np.array([wl["root"]], dtype=np.uint64))
    t.free()


def test_circuit_fingerprint(ctx, m):
    """get_circuit_fingerprint_generic (pm_core.rs:18-42) = hash_no_pad(constants_sigmas_cap || circuit_digest)"""
    cols = [rand_felts(0xF1F0 + c, 1 << 6) for c in range(9)]
    b = m.PolynomialBatch.from_values(ctx, cols, 3, False, 4)
    digest = [11, 12, 13, 14]
    ref = O.batch_from_values(cols, 3, 4)
    want = O.hash_no_pad(list(ref["cap"].reshape(-1)) + digest)
    assert m.circuit_fingerprint(ctx, b, digest).tolist() == want.tolist()
    b.free()


def test_hash_no_pad_lengths(ctx):
    for n in (0, 1, 4, 5, 7, 8, 9, 16, 17, 135):
        x = rand_felts(n + 1, n, canonical=False)
        assert (ctx.hash_no_pad(x) == O.hash_no_pad(x)).all(), n


def test_k3_stored_proof_paths_on_gpu(ctx, m, golden_dir):
    """qbench_data/example.bin: leaf hashing (135/20/16/32-wide) and path climbing on the GPU must reach the
    caps stored in the proofs."""
    idx = json.load(open(os.path.join(golden_dir, "example_proofs.json")))["proofs"]
    blob = open(os.path.join(golden_dir, "example_proofs.bin"), "rb").read()
    p = parse_proof(blob[idx[0]["offset"] : idx[0]["offset"] + idx[0]["len"]])
    caps = [None, p["wires_cap"], p["zs_pp_cap"], p["quotient_cap"]]
    rounds = p["query_rounds"]
    x_indices = [O.merkle_find_index(r["initial"][3][0], r["initial"][3][1], caps[3]) for r in rounds]
    for t in (1, 2, 3):
        leaves = np.stack([r["initial"][t][0] for r in rounds] + [rounds[0]["initial"][t][0]] * 4)  # pad 28 -> 32
        tree = m.MerkleTree.new(ctx, leaves, 5)  # cap_height = log2(32): the cap is the leaf digests
        cur = tree.cap[:28].copy()
        idxs = np.array(x_indices)
        for lvl in range(11):
            sib = np.stack([r["initial"][t][1][lvl] for r in rounds])
            bit = (idxs >> lvl) & 1
            left = np.where(bit[:, None] == 1, sib, cur)
            right = np.where(bit[:, None] == 1, cur, sib)
            cur = ctx.two_to_one(left, right)
        for k in range(28):
            assert (cur[k] == caps[t][x_indices[k] >> 11]).all()
    # FRI layer 0 leaves: 16 extension evaluations = 32 felts
    leaves = np.stack([r["steps"][0][0].reshape(-1) for r in rounds] + [rounds[0]["steps"][0][0].reshape(-1)] * 4)
    tree = m.MerkleTree.new(ctx, leaves, 5)
    cur = tree.cap[:28].copy()
    idxs = np.array(x_indices) >> 4
    for lvl in range(7):
        sib = np.stack([r["steps"][0][1][lvl] for r in rounds])
        bit = (idxs >> lvl) & 1
        cur = ctx.two_to_one(np.where(bit[:, None] == 1, sib, cur), np.where(bit[:, None] == 1, cur, sib))
    for k in range(28):
        assert (cur[k] == p["commit_phase_merkle_caps"][0][idxs[k] >> 7]).all()


# ----------------------------------------------------------------------------- MerkleTree::new
@pytest.mark.parametrize("n_leaves,leaf_len,cap_height", [
    (1, 5, 0), (2, 3, 0), (2, 3, 1), (16, 4, 2), (16, 5, 0), (64, 7, 3), (64, 8, 6), (128, 9, 4),
    (256, 32, 4), (512, 135, 4), (1024, 20, 0), (2048, 1, 4), (4096, 16, 4)])
def test_merkle_tree_new(ctx, m, n_leaves, leaf_len, cap_height):
    leaves = rand_felts(n_leaves * 131 + leaf_len, (n_leaves, leaf_len), canonical=False)
    dg, cap = O.merkle_tree_new(leaves, cap_height)
    t = m.MerkleTree.new(ctx, leaves, cap_height)
    assert (t.cap == cap).all()
    assert (t.digests == dg).all()  # plonky2's interleaved layout
    for i in sorted({0, n_leaves - 1, n_leaves // 2, (n_leaves * 5) // 7}):
        sib = t.prove(i)
        assert (sib == O.merkle_prove(dg, n_leaves, cap_height, i)).all()
        assert O.merkle_verify(leaves[i], i, sib, cap)
        assert (t.get(i) == leaves[i]).all()


def test_merkle_errors(ctx, m):
    with pytest.raises(m.P2BError):
        m.MerkleTree.new(ctx, np.zeros((3, 4), np.uint64), 0)  # not a power of two
    with pytest.raises(m.P2BError):
        m.MerkleTree.new(ctx, np.zeros((4, 4), np.uint64), 3)  # cap above the leaves
    t = m.MerkleTree.new(ctx, np.zeros((4, 4), np.uint64), 1)
    with pytest.raises(m.P2BError):
        t.prove(4)


# ----------------------------------------------------------------------------- PolynomialBatch
def _check_batch(ctx, m, cols, rate_bits, cap_height, from_values, sample_only=False):
    n = cols[0].size
    log_n = n.bit_length() - 1
    N = n << rate_bits
    ref = (O.batch_from_values if from_values else O.batch_from_coeffs)(cols, rate_bits, cap_height)
    mk = m.PolynomialBatch.from_values if from_values else m.PolynomialBatch.from_coeffs
    b = mk(ctx, cols, rate_bits, False, cap_height)
    assert (b.cap == ref["cap"]).all()
    if from_values:
        for c in sorted({0, len(cols) - 1, len(cols) // 2}):
            assert (b.coeffs(c) == ref["coeffs"][c]).all(), c
    else:
        assert (b.coeffs(0) == (cols[0] % np.uint64(P))).all()
    if not sample_only:
        assert (b.leaves() == ref["leaves"]).all()
        assert (b.merkle_tree.digests == ref["digests"]).all()
    for j in sorted({0, 1 % N, N - 1, N // 3, (N * 5) // 7}):
        assert (b.leaf(j) == ref["leaves"][j]).all(), j
        assert (b.get_lde_values(bitrev(j, log_n + rate_bits), 1) == ref["leaves"][j]).all()
        sib = b.merkle_tree.prove(j)
        assert (sib == O.merkle_prove(ref["digests"], N, cap_height, j)).all()
        assert O.merkle_verify(ref["leaves"][j], j, sib, ref["cap"])
    b.free()


@pytest.mark.parametrize("log_n,n_cols,rate_bits,cap_height", [
    (0, 3, 3, 0), (1, 2, 3, 1), (2, 5, 3, 4), (3, 9, 1, 0), (5, 20, 3, 4), (8, 16, 3, 4), (10, 135, 3, 4),
    (12, 135, 3, 4), (12, 20, 3, 4), (12, 16, 3, 4), (12, 85, 3, 4), (12, 1, 0, 0), (13, 17, 3, 4),
    (14, 20, 3, 4), (14, 37, 3, 4), (15, 8, 2, 4), (16, 4, 3, 4), (17, 2, 3, 4), (18, 2, 2, 4), (19, 1, 1, 4)])
def test_batch_from_values(ctx, m, log_n, n_cols, rate_bits, cap_height):
    cols = [rand_felts(0x5EED0001 + c, 1 << log_n, canonical=(c % 3 != 0)) for c in range(n_cols)]
    _check_batch(ctx, m, cols, rate_bits, cap_height, True)


@pytest.mark.parametrize("log_n,n_cols,rate_bits,cap_height", [
    (4, 16, 3, 4), (12, 16, 3, 4), (13, 16, 3, 4), (14, 33, 2, 4), (16, 2, 3, 4)])
def test_batch_from_coeffs(ctx, m, log_n, n_cols, rate_bits, cap_height):
    cols = [rand_felts(0xC0EFF + c, 1 << log_n, canonical=(c % 2 == 0)) for c in range(n_cols)]
    _check_batch(ctx, m, cols, rate_bits, cap_height, False)


def test_batch_random_shapes(ctx, m):
    """seeded random shapes around the kernel-selection boundaries (generic smem NTT below 2^12, radix-16 path with
    1..4 strided bits above it, pipelined upload for >= 32 columns at >= 2^14 rows), values and coefficients"""
    rng = np.random.default_rng(20261018)
    for case in range(10):
        log_n = int(rng.integers(9, 16))
        n_cols = int(rng.choice([1, 2, 7, 19, 33, 48]))
        rate_bits = int(rng.integers(0, 4))
        cap_height = int(rng.integers(0, min(5, log_n + rate_bits + 1)))
        cols = [rand_felts(7000 + 100 * case + c, 1 << log_n, canonical=bool((c + case) % 2)) for c in range(n_cols)]
        _check_batch(ctx, m, cols, rate_bits, cap_height, from_values=bool(case % 2 == 0), sample_only=log_n + rate_bits > 15)


def test_pinned_and_pageable_inputs_agree(ctx, m):
    """pinned host columns go by direct DMA, pageable ones through the staging double buffer (incl. a column longer
    than one 8 MiB half): same batch"""
    log_n, n_cols = 21, 2  # 16 MiB per column
    pageable = [rand_felts(31 + c, 1 << log_n) for c in range(n_cols)]
    pinned = ctx.pinned_empty((n_cols, 1 << log_n))
    for c in range(n_cols):
        pinned[c] = pageable[c]
    a = m.PolynomialBatch.from_values(ctx, pageable, 1, False, 3)
    b = m.PolynomialBatch.from_values(ctx, [pinned[c] for c in range(n_cols)], 1, False, 3)
    assert (a.cap == b.cap).all()
    assert (a.coeffs(1) == b.coeffs(1)).all()
    a.free()
    b.free()


def test_separately_allocated_and_mixed_host_columns(ctx, m):
    """plonky2's witness is one allocation per column: consecutive pageable columns are packed into the staging buffer
    wherever they lie; pinned columns in between go by direct DMA.  Same batch as from one contiguous matrix."""
    log_n, n_cols = 10, 37
    base = rand_felts(77, (n_cols, 1 << log_n))
    keep = [np.zeros(5 + 3 * c, np.uint64) for c in range(n_cols)]  # spread the allocations
    separate = [base[c].copy() for c in range(n_cols)]
    pinned = ctx.pinned_empty((4, 1 << log_n))
    mixed = list(separate)
    for k, c in enumerate((0, 9, 10, 36)):
        pinned[k] = base[c]
        mixed[c] = pinned[k]
    ref = m.PolynomialBatch.from_values(ctx, base, 3, False, 4)
    for cols in (separate, mixed):
        b = m.PolynomialBatch.from_values(ctx, cols, 3, False, 4)
        assert (b.cap == ref.cap).all()
        for c in (0, 8, 9, 10, 11, 36):
            assert (b.coeffs(c) == ref.coeffs(c)).all()
        b.free()
    ref.free()
    del keep


def test_batch_config1_shape_bit_exact(ctx, m):
    """BASELINE.json configs[1]: 2^16 rows x 135 wire columns, rate_bits 3, cap_height 4 — the full commit
    against the oracle (coefficients, sampled leaves, paths, cap)."""
    cols = [rand_felts(0x5EED0001 + c, 1 << 16) for c in range(135)]
    _check_batch(ctx, m, cols, 3, 4, True, sample_only=True)


def test_batch_errors(ctx, m):
    with pytest.raises(m.P2BError):
        m.PolynomialBatch.from_values(ctx, [np.zeros(8, np.uint64)], 3, True, 4)  # blinding unsupported
    with pytest.raises(m.P2BError):
        m.PolynomialBatch.from_values(ctx, [np.zeros(2, np.uint64)], 1, False, 4)  # cap above the leaves
    with pytest.raises(ValueError):
        m.PolynomialBatch.from_values(ctx, [np.zeros(6, np.uint64)], 3, False, 1)
    with pytest.raises(ValueError):
        m.PolynomialBatch.from_values(ctx, [], 3, False, 1)
    # the context stays usable after argument errors
    assert (ctx.hash_no_pad([1, 2, 3, 4, 5]) == O.hash_no_pad([1, 2, 3, 4, 5])).all()


def _horner(coeffs, x):
    acc = 0
    for c in reversed([int(c) for c in coeffs]):
        acc = (acc * x + c) % P
    return acc


def test_large_batch_properties_2p20(ctx, m):
    """2^20 rows (the M2 shape's row count) x 6 columns: too big for a full oracle run in seconds, so use
    size-independent properties: ifft/fft round trip at sampled points, linearity, Merkle paths."""
    log_n, rate = 20, 3
    n = 1 << log_n
    a = [rand_felts(1 + c, n) for c in range(3)]
    bcols = [rand_felts(77 + c, n) for c in range(3)]
    s = [((x.astype(object) + y.astype(object)) % P).astype(np.uint64) for x, y in zip(a, bcols)]
    batch = m.PolynomialBatch.from_values(ctx, a + bcols, rate, False, 4)
    bsum = m.PolynomialBatch.from_values(ctx, s, rate, False, 4)
    log_N = log_n + rate
    w = O.root_of_unity(log_N)
    cap = batch.cap
    for j in (0, 12345, (1 << log_N) - 1, 5 << 19):
        leaf = batch.leaf(j)
        lsum = bsum.leaf(j)
        for c in range(3):  # linearity of the whole pipeline
            assert int(lsum[c]) == (int(leaf[c]) + int(leaf[3 + c])) % P
        assert O.merkle_verify(leaf, j, batch.merkle_tree.prove(j), cap)
    # LDE restricted to the subgroup coset: evaluate the coefficients directly at one point
    coeffs0 = batch.coeffs(0)
    j = 987654
    x = 7 * pow(w, bitrev(j, log_N), P) % P
    assert int(batch.leaf(j)[0]) == _horner(coeffs0, x)
    # ifft really inverts: coefficients evaluated at w_n^i give back value i
    wn = O.root_of_unity(log_n)
    for i in (0, 1, 54321):
        assert _horner(coeffs0, pow(wn, i, P)) == int(a[0][i])
    batch.free()
    bsum.free()


def test_m2_shape_2p20x135_against_oracle(ctx, m):
    """The M2 shape of BASELINE.json's metric (2^20 rows x 135 columns, rate 8, cap 4) at full size: three sampled
    columns are compared with the oracle ELEMENT FOR ELEMENT — 2^20 coefficients and 2^23 LDE values each — and eight
    sampled leaves open against the GPU cap through the oracle's leaf hashing and Merkle verification."""
    import torch

    log_n, rate, n_cols = 20, 3, 135
    n, log_N = 1 << log_n, log_n + rate
    g = torch.Generator(device="cuda")
    g.manual_seed(135)
    vals = torch.empty((n_cols, n), dtype=torch.int64, device="cuda")
    vals.random_(0, 2**62, generator=g)
    vals[67] -= 2**62  # a column with words >= 2^63, incl. non-canonical ones
    vals[67, 5] = -1
    torch.cuda.synchronize()
    batch = m.PolynomialBatch.from_values_device(ctx, vals.data_ptr(), n_cols, log_n, rate, 4)
    cap = batch.cap
    rev = np.arange(1 << log_N, dtype=np.uint64)
    out = np.zeros_like(rev)
    for b in range(log_N):  # vectorised reverse_index_bits
        out |= ((rev >> np.uint64(b)) & np.uint64(1)) << np.uint64(log_N - 1 - b)
    ref_lde = {}
    for c in (0, 67, 134):
        col = vals[c].cpu().numpy().view(np.uint64)
        coeffs = O.ifft(col)
        assert (batch.coeffs(c) == coeffs).all(), "coefficients of column %d" % c
        padded = np.zeros(1 << log_N, np.uint64)
        padded[:n] = coeffs
        ref_lde[c] = O.coset_fft(padded, 7)[out]  # leaf j holds LDE row bitrev(j)
        assert (batch.lde_col(c) == ref_lde[c]).all(), "LDE of column %d" % c
    for j in (0, 1, (1 << log_N) - 1, 5 << 19, 1234567, 7654321, 1 << 22, (1 << 22) + 1):
        leaf = batch.leaf(j)
        for c, lde in ref_lde.items():
            assert leaf[c] == lde[j]
        assert O.merkle_verify(leaf, j, batch.merkle_tree.prove(j), cap)
    batch.free()
    del vals
    torch.cuda.empty_cache()


def test_config3_commit_2p20x400_properties(ctx, m):
    """BASELINE.json configs[2]: the 2^20 rows x 400 columns commit at full size (33.6 GB on the device).  No oracle
    run at this size; size-independent properties instead: sampled leaves open against the cap (oracle hashing of
    the 400-wide leaf, 19 siblings), the coefficients interpolate the input values, the LDE is their evaluation on
    the shifted coset, and a column's LDE does not depend on the batch around it."""
    import torch

    log_n, rate, n_cols = 20, 3, 400
    n, log_N = 1 << log_n, log_n + rate
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 60 << 30:
        pytest.skip("needs ~45 GB of free device memory")
    g = torch.Generator(device="cuda")
    g.manual_seed(400)
    vals = torch.empty((n_cols, n), dtype=torch.int64, device="cuda")
    vals.random_(0, 2**62, generator=g)
    vals[3] -= 2**62  # a column of u64 values >= 2^63, with two explicit non-canonical ones (2^64 - 1 and p itself)
    vals[3, 0] = -1
    vals[3, 1] = -(2**32) + 1
    torch.cuda.synchronize()  # the library works on its own stream
    batch = m.PolynomialBatch.from_values_device(ctx, vals.data_ptr(), n_cols, log_n, rate, 4)
    cap = batch.cap
    assert cap.shape == (16, 4)
    w = O.root_of_unity(log_N)
    wn = O.root_of_unity(log_n)
    cols = (0, 3, 399)
    coeffs = {c: batch.coeffs(c) for c in cols}
    for c in cols:
        assert (coeffs[c] < np.uint64(P)).all()
    samples = (0, 1, (1 << log_N) - 1, 5 << 19, 1234567)
    for j in samples:
        leaf = batch.leaf(j)
        assert leaf.shape == (n_cols,) and (leaf < np.uint64(P)).all()
        sib = batch.merkle_tree.prove(j)
        assert sib.shape == (log_N - 4, 4)
        assert O.merkle_verify(leaf, j, sib, cap)
        bad = leaf.copy()
        bad[217] ^= np.uint64(1)
        assert not O.merkle_verify(bad, j, sib, cap)
    for c in cols:
        col_vals = vals[c].cpu().numpy().view(np.uint64)
        for i in (0, 1, 777777):  # ifft really inverts
            assert _horner(coeffs[c], pow(wn, i, P)) == int(col_vals[i]) % P
        j = samples[-1]  # leaf j = the evaluation at 7 w^bitrev(j)
        assert int(batch.leaf(j)[c]) == _horner(coeffs[c], 7 * pow(w, bitrev(j, log_N), P) % P)
    sub = torch.stack([vals[c] for c in cols]).contiguous()
    torch.cuda.synchronize()
    small = m.PolynomialBatch.from_values_device(ctx, sub.data_ptr(), len(cols), log_n, rate, 4)
    for j in samples:
        big_leaf, small_leaf = batch.leaf(j), small.leaf(j)
        for k, c in enumerate(cols):
            assert big_leaf[c] == small_leaf[k]
    small.free()
    batch.free()
    del vals, sub
    torch.cuda.empty_cache()


def test_config3_fri_commit_2p23_bit_exact(ctx, m):
    """BASELINE.json configs[2], FRI part: fri_committed_trees over an extension polynomial of 8n = 2^23 points with
    reduction_arity_bits [4,4,4,4] — the whole commit phase against the oracle (about ten seconds of CPU): caps,
    final polynomial, transcript, sampled leaves and paths of every layer."""
    log_n, rate_bits, cap_height, arity_bits = 20, 3, 4, [4, 4, 4, 4]
    n = 1 << log_n
    N = n << rate_bits
    coeffs = np.zeros((N, 2), np.uint64)
    coeffs[:n] = rand_felts(2023, (n, 2))
    values = O.ext_coset_fft(coeffs, 7)
    oc = O.Challenger()
    oc.observe([1, 2, 3, 4])
    ref = O.fri_committed_trees(coeffs, values, arity_bits, oc, rate_bits, cap_height)
    gc = m.Challenger(ctx)
    gc.observe_elements([1, 2, 3, 4])
    trees, final = m.fri_committed_trees(ctx, coeffs, values, gc, arity_bits, rate_bits, cap_height)
    assert final.shape == (16, 2) and (final == ref["final_poly"]).all()
    assert len(trees) == 4
    for l, t in enumerate(trees):
        assert t.n_leaves == N >> (4 * (l + 1))
        assert (t.cap == ref["caps"][l]).all(), l
        for j in sorted({0, t.n_leaves - 1, t.n_leaves // 3}):
            assert (t.get(j) == ref["leaves"][l][j]).all()
            assert (t.prove(j) == O.merkle_prove(ref["digests"][l], t.n_leaves, cap_height, j)).all()
    assert gc.export_state()[:12].tolist() == oc.state_words()[0]


# ----------------------------------------------------------------------------- Challenger / FRI
def test_challenger_matches_oracle(ctx, m):
    rng = np.random.default_rng(3)
    g, o = m.Challenger(ctx), O.Challenger()
    for step in range(40):
        k = int(rng.integers(0, 20))
        e = rand_felts(1000 + step, k, canonical=(step % 2 == 0))
        g.observe_elements(e)
        o.observe(e)
        if step % 3 != 1:
            q = int(rng.integers(1, 12))
            assert g.get_n_challenges(q) == o.get_n(q)
    st, inb = o.state_words()
    ex = g.export_state()
    assert ex[:12].tolist() == st and int(ex[12]) == len(inb) and ex[13:13 + len(inb)].tolist() == inb
    g2 = m.Challenger(ctx)
    g2.import_state(ex)
    assert g2.get_n_challenges(9) == g.get_n_challenges(9)


@pytest.mark.parametrize("log_n,arity_bits,rate_bits,cap_height", [
    (5, [4], 3, 0), (7, [4, 3], 3, 1), (9, [4, 4], 3, 4), (12, [4, 4], 3, 4), (12, [1, 2, 3, 4], 3, 2),
    (13, [4, 4], 3, 4), (14, [4, 4, 4], 2, 4), (16, [4, 4, 4], 3, 4)])
def test_fri_committed_trees(ctx, m, log_n, arity_bits, rate_bits, cap_height):
    n = 1 << log_n
    N = n << rate_bits
    coeffs = np.zeros((N, 2), np.uint64)
    coeffs[:n] = rand_felts(900 + log_n, (n, 2))
    values = O.ext_coset_fft(coeffs, 7)
    seed = rand_felts(4, 5)
    oc = O.Challenger()
    oc.observe(seed)
    ref = O.fri_committed_trees(coeffs, values, arity_bits, oc, rate_bits, cap_height)
    gc = m.Challenger(ctx)
    gc.observe_elements(seed)
    trees, final = m.fri_committed_trees(ctx, coeffs, values, gc, arity_bits, rate_bits, cap_height)
    assert (final == ref["final_poly"]).all()
    for l, t in enumerate(trees):
        assert (t.cap == ref["caps"][l]).all(), l
        assert (t.digests == ref["digests"][l][: t.digests.shape[0]]).all(), l
        nl = t.n_leaves
        for j in sorted({0, nl - 1, nl // 3}):
            assert (t.get(j) == ref["leaves"][l][j]).all()
            assert (t.prove(j) == O.merkle_prove(ref["digests"][l], nl, cap_height, j)).all()
    # transcripts agree after the commit phase, and so does the proof of work (minimal witness)
    assert gc.export_state()[:12].tolist() == oc.state_words()[0]
    wg = m.fri_proof_of_work(ctx, gc, 12)
    wo = O.fri_proof_of_work(oc, 12)
    assert wg == wo
    assert gc.get_n_challenges(4) == oc.get_n(4)


def test_pow_16_bits_is_minimal(ctx, m):
    gc, oc = m.Challenger(ctx), O.Challenger()
    gc.observe_elements([5, 6, 7, 8, 9])
    oc.observe([5, 6, 7, 8, 9])
    base = oc.clone()
    w = m.fri_proof_of_work(ctx, gc, 16)
    assert O.fri_pow_check(base, w, 16) == 1
    assert w == O.fri_proof_of_work(oc, 16)
