"""GPU: the complete prover flow behind the C ABI (wires commit -> Z / partial products -> quotient -> openings ->
prove_openings: final polynomial, FRI commit phase, proof of work, query rounds) against the oracle-built proof
(tests/verifier_ref.py::oracle_prove), word for word, and through the restated plonky2 verifier."""
import os

import numpy as np
import pytest

import p2oracle as O
import plonk_ref as R
import verifier_ref as V
from util import rand_felts
from test_plonk_oracle import ALL_GATES, CITY_GATES, CITY_GROUPS, RECURSION_GATES, RECURSION_GROUPS
from test_prove_oracle import FP_SMALL, make_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL_GROUPS = [(0, 4), (4, 5), (5, 8), (8, 10)]
# the FRI / Plonk parameters of every stored City Rollup proof (city_common_circuit/src/circuits/zk_signature2/mod.rs:33-57)
FP_CITY = dict(rate_bits=3, cap_height=4, proof_of_work_bits=16, num_query_rounds=28, reduction_arity_bits=[4, 4])


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


def gpu_prove(ctx, m, circ, digest, pis, fp):
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"],
                                       keep_values=True)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"],
                         fp["reduction_arity_bits"])
    proof = m.prove(ctx, cd, cs, digest, circ.wire_values(), pis, params)
    native = m.prove_native(ctx, cd, cs, digest, circ.wire_values(), pis, params)
    assert V.proofs_equal(native, proof) is None, "p2b_prove differs from the stage-by-stage flow: %s" % V.proofs_equal(native, proof)
    cs_cap = cs.cap
    cs.free()
    cd.free()
    return proof, cs_cap


@pytest.mark.parametrize("degree_bits,gates,groups,seed,fp", [
    (6, ALL_GATES, FULL_GROUPS, 51, FP_SMALL),
    (5, ALL_GATES[:5], [(0, 4), (4, 5)], 52, dict(FP_SMALL, cap_height=0, reduction_arity_bits=[1, 2, 1], num_query_rounds=3)),
    (9, ALL_GATES, FULL_GROUPS, 53, dict(FP_SMALL, cap_height=4, reduction_arity_bits=[4, 4], proof_of_work_bits=10)),
    (12, ALL_GATES, FULL_GROUPS, 54, FP_CITY),
    (10, RECURSION_GATES, RECURSION_GROUPS, 55, FP_CITY),
])
def test_gpu_proof_equals_oracle_proof_and_verifies(ctx, m, degree_bits, gates, groups, seed, fp):
    circ, digest, pis = make_case(degree_bits, gates, groups, seed)
    got, cs_cap = gpu_prove(ctx, m, circ, digest, pis, fp)
    ref, ref_cs_cap = V.oracle_prove(circ, digest, pis, fp)
    assert (cs_cap == ref_cs_cap).all()
    assert V.proofs_equal(got, ref) is None, V.proofs_equal(got, ref)
    assert V.verify(circ, cs_cap, digest, got, fp)


@pytest.mark.parametrize("name,gates,groups,seed", [
    ("recursion", RECURSION_GATES, RECURSION_GROUPS, 61),  # the gate set of the proofs stored in qbench_data/example.bin
    ("city", CITY_GATES, CITY_GROUPS, 62),                 # add_city_common_gates + the in-tree u32 gates: bench.py's M1 circuit
])
def test_city_shape_proof_equals_c_oracle(ctx, m, name, gates, groups, seed):
    """EXACTLY the M1 case of bench.py — 2^12 rows x 135 wires, 28 queries, 16-bit PoW, arities [4, 4] — through p2b_prove
    (host witness) and p2b_prove_dev (witness in HBM), word for word against the C oracle prover (oracle/prove.c), and
    through the restated verifier."""
    import torch

    circ, digest, pis = make_case(12, gates, groups, seed)
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
    params = m.FriParams(3, 4, 16, 28, [4, 4])
    wv = circ.wire_values()
    got = m.prove_native(ctx, cd, cs, digest, wv, pis, params, raw=True)
    dev = torch.from_numpy(np.stack(wv).view(np.int64)).cuda()
    torch.cuda.synchronize()
    got_dev = m.prove_native_device(ctx, cd, cs, digest, dev.data_ptr(), pis, params)
    want, ref_cs_cap = V.oracle_prove_c(circ, digest, pis, FP_CITY)
    assert (cs.cap == ref_cs_cap).all()
    assert got.shape == want.shape
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, "%s: first differing proof word %d of %d" % (name, bad[0], want.size)
    assert (got_dev == want).all(), "p2b_prove_dev differs"
    assert V.verify(circ, cs.cap, digest, V.parse_proof(circ, FP_CITY, got, len(pis)), FP_CITY)
    cs.free()
    cd.free()


def _random_witness(circ, seed):
    """any words at all: the prover does not check satisfaction, and the oracle prover computes the same (invalid) proof"""
    rng = np.random.default_rng(seed)
    return [rng.integers(0, R.P, circ.n, dtype=np.uint64) for _ in range(circ.num_wires)]


@pytest.mark.parametrize("degree_bits,fp", [(8, dict(FP_SMALL, cap_height=3, proof_of_work_bits=8)), (12, FP_CITY)])
def test_replayed_plans_equal_the_oracle(m, degree_bits, fp):
    """From its second proof of a shape on, a context replays a captured CUDA graph (p2b_plan_info says so).  Different
    witnesses, public inputs and circuit digests through the SAME plan — host witness (pageable columns, one pageable
    matrix, a pinned matrix), device witness, and the submit / collect form — must each equal the C oracle's proof
    word for word; two circuits alternate on the context to exercise two plans side by side."""
    import torch

    c = m.Context(0)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], fp["reduction_arity_bits"])
    cases = []
    for k, (gates, groups) in enumerate(((RECURSION_GATES, RECURSION_GROUPS), (CITY_GATES, CITY_GROUPS))):
        circ, digest, pis = make_case(degree_bits, gates, groups, 70 + k)
        cd = m.CircuitData(c, circ.desc())
        cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"], keep_values=True)
        pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), fp)
        cases.append((circ, cd, cs, pd))
    pinned = c.pinned_empty((cases[0][0].num_wires, cases[0][0].n))
    for it in range(5):
        for k, (circ, cd, cs, pd) in enumerate(cases):
            wv = circ.wire_values() if it == 0 else _random_witness(circ, 1000 * k + it)
            digest = [it + 1, 2 * k + 5, 0xFFFFFFFF00000000 - it, 9]
            pis = [it, k, 3, 0xFFFFFFFF00000001 + it]  # incl. a non-canonical word
            want = pd.prove(digest, wv, pis)
            mode = ("cols", "matrix", "pinned", "submit", "cols")[it]
            if mode == "cols":
                got = m.prove_native(c, cd, cs, digest, wv, pis, params, raw=True)
            elif mode == "matrix":
                got = m.prove_native(c, cd, cs, digest, np.stack(wv), pis, params, raw=True)
            elif mode == "pinned":
                pinned[:] = np.stack(wv)
                got = m.prove_native(c, cd, cs, digest, pinned, pis, params, raw=True)
            else:
                scratch = np.stack(wv)
                n_words = m.prove_submit(c, cd, cs, digest, scratch, pis, params)
                scratch[:] = 0  # the witness buffer belongs to the caller again as soon as submit returns
                got = m.prove_collect(c, n_words)
            assert (got == want).all(), "iteration %d circuit %d (%s): first differing word %d" % (it, k, mode, np.nonzero(got != want)[0][0])
            dev = torch.from_numpy(np.stack(wv).view(np.int64)).cuda()
            torch.cuda.synchronize()
            got_dev = m.prove_native_device(c, cd, cs, digest, dev.data_ptr(), pis, params)
            assert (got_dev == want).all(), "iteration %d circuit %d: p2b_prove_dev differs" % (it, k)
    info = c.plan_info()
    if os.environ.get("P2B_GRAPH", "1") != "0":
        assert info["failed"] == 0, info
        assert info["ready"] == 4, info  # 2 circuits x (host, device) witness
        assert 0 < info["kernels_per_launch"] <= 72, info  # 58 with the Merkle trees fused from 2^15 digests; 70 since they are fused from 2^11 (+9 % proofs/s, p2b.cu build_levels)
    for circ, cd, cs, pd in cases:
        pd.free()
        cs.free()
        cd.free()
    c.close()


def test_one_thread_drives_several_contexts(m):
    """p2b_prove_submit / p2b_prove_collect: ONE host thread keeps three contexts busy (each with a proof in flight) and
    shares ONE constants|sigmas batch between them (p2b_batch_attach); every proof equals the oracle's."""
    fp = dict(FP_SMALL, cap_height=3, proof_of_work_bits=8)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], fp["reduction_arity_bits"])
    circ, digest, pis = make_case(9, CITY_GATES, CITY_GROUPS, 81)
    ctxs = [m.Context(0) for _ in range(3)]
    owner = m.PolynomialBatch.from_values(ctxs[0], circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"], keep_values=True)
    views = [owner] + [owner.attach(c) for c in ctxs[1:]]
    cds = [m.CircuitData(c, circ.desc()) for c in ctxs]
    pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), fp)
    for rnd in range(4):
        wvs = [circ.wire_values() if (rnd + i) % 3 == 0 else _random_witness(circ, 50 * rnd + i) for i in range(3)]
        n_words = [m.prove_submit(c, cd, v, digest, np.stack(w), pis, params) for c, cd, v, w in zip(ctxs, cds, views, wvs)]
        with pytest.raises(m.P2BError):  # one pending proof per context
            m.prove_submit(ctxs[0], cds[0], views[0], digest, np.stack(wvs[0]), pis, params)
        for c, nw, w in zip(ctxs, n_words, wvs):
            got = m.prove_collect(c, nw)
            assert (got == pd.prove(digest, w, pis)).all()
    with pytest.raises(m.P2BError):
        m.prove_collect(ctxs[1], n_words[1])  # nothing pending
    for v in views[1:]:
        v.free()
    for cd in cds:
        cd.free()
    owner.free()
    pd.free()
    for c in ctxs:
        c.close()


def test_submit_nowait_with_pinned_witness(m):
    """p2b_prove_submit_nowait / p2b_prove_upload_poll: the driver-thread form (no wait for the upload inside submit).  Two
    contexts with pinned witness matrices; the upload poll turns true, the buffer is refilled only after that, and every
    proof equals the oracle's."""
    fp = dict(FP_SMALL, cap_height=3, proof_of_work_bits=8)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], fp["reduction_arity_bits"])
    circ, digest, pis = make_case(9, CITY_GATES, CITY_GROUPS, 83)
    ctxs = [m.Context(0) for _ in range(2)]
    cds = [m.CircuitData(c, circ.desc()) for c in ctxs]
    css = [m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"], keep_values=True)
           for c in ctxs]
    pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), fp)
    shape = np.stack(circ.wire_values()).shape
    bufs = [c.pinned_empty(shape) for c in ctxs]
    assert m.prove_upload_poll(ctxs[0])  # nothing pending: nothing reads host memory
    for rnd in range(3):
        wvs = [circ.wire_values() if (rnd + i) % 2 == 0 else _random_witness(circ, 70 * rnd + i) for i in range(2)]
        for b, w in zip(bufs, wvs):
            b[:] = np.stack(w)
        n_words = [m.prove_submit(c, cd, cs, digest, b, pis, params, wait_upload=False) for c, cd, cs, b in zip(ctxs, cds, css, bufs)]
        for c in ctxs:
            while not m.prove_upload_poll(c):
                pass
        for b in bufs:
            b[:] = 0  # the witness has been read: the worker may build the next one in place
        for c, nw, w in zip(ctxs, n_words, wvs):
            assert (m.prove_collect(c, nw) == pd.prove(digest, w, pis)).all()
    for cs in css:
        cs.free()
    for cd in cds:
        cd.free()
    pd.free()
    for c in ctxs:
        c.close()


def test_latency_mode_proofs_equal_throughput_mode(m):
    """p2b_set_latency_mode: other launch configurations (trees fused from 2^15 digests, proof-of-work search on every SM),
    the same proof words — eagerly, from a captured plan, and after switching back (the switch drops the plans)"""
    fp = dict(FP_SMALL, cap_height=3, proof_of_work_bits=10)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], fp["reduction_arity_bits"])
    circ, digest, pis = make_case(10, CITY_GATES, CITY_GROUPS, 85)
    c = m.Context(0)
    cd = m.CircuitData(c, circ.desc())
    cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"], keep_values=True)
    pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), fp)
    want = pd.prove(digest, circ.wire_values(), pis)
    for mode in (True, True, False, True):
        c.set_latency_mode(mode)
        for _ in range(3):  # eager, capture, replay
            got = m.prove_native(c, cd, cs, digest, circ.wire_values(), pis, params, raw=True)
            assert (got == want).all(), mode
    n = m.prove_submit(c, cd, cs, digest, np.stack(circ.wire_values()), pis, params)
    with pytest.raises(m.P2BError):
        c.set_latency_mode(False)  # a proof is pending
    assert (m.prove_collect(c, n) == want).all()
    cs.free()
    cd.free()
    pd.free()
    c.close()


def test_constants_sigmas_export_import_round_trip(ctx, m):
    """p2b_batch_export / p2b_batch_import: the imported batch (coefficients + kept values + cap; LDE
    and tree recomputed on the device) proves the same proof; a flipped byte is rejected through the cap check."""
    fp = dict(FP_SMALL, cap_height=3, proof_of_work_bits=8)
    params = m.FriParams(fp["rate_bits"], fp["cap_height"], fp["proof_of_work_bits"], fp["num_query_rounds"], fp["reduction_arity_bits"])
    circ, digest, pis = make_case(8, RECURSION_GATES, RECURSION_GROUPS, 91)
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), fp["rate_bits"], False, fp["cap_height"], keep_values=True)
    blob = cs.export()
    c2 = m.Context(0)
    cs2 = m.PolynomialBatch.import_(c2, blob)
    assert (cs2.cap == cs.cap).all() and cs2.n_cols == cs.n_cols
    for j in (0, 77, (1 << 11) - 1):
        assert (cs2.leaf(j) == cs.leaf(j)).all()
        assert (cs2.merkle_tree.prove(j) == cs.merkle_tree.prove(j)).all()
    assert (cs2.values(3) == cs.values(3)).all()
    cd2 = m.CircuitData(c2, circ.desc())
    a = m.prove_native(ctx, cd, cs, digest, circ.wire_values(), pis, params, raw=True)
    b = m.prove_native(c2, cd2, cs2, digest, circ.wire_values(), pis, params, raw=True)
    assert (a == b).all()
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 1
    with pytest.raises(m.P2BError) as e:
        m.PolynomialBatch.import_(c2, bytes(bad))
    assert "cap" in str(e.value)
    with pytest.raises(m.P2BError):
        m.PolynomialBatch.import_(c2, blob[:-8])
    cs2.free()
    cd2.free()
    c2.close()
    cs.free()
    cd.free()


def test_pinned_input_may_be_refilled_after_return(ctx, m):
    """host-input entry points return only after their input has been read: overwriting a pinned witness / value buffer
    right after the call must not change the result (the DMA used to be still in flight)"""
    log_n, n_cols = 14, 40  # the pipelined upload path (>= 32 columns of >= 2^14 rows) and the plain one
    for ln, nc in ((log_n, n_cols), (10, 20)):
        src = ctx.pinned_empty((nc, 1 << ln))
        orig = np.stack([rand_felts(900 + c, 1 << ln) for c in range(nc)])
        src[:] = orig
        b = m.PolynomialBatch.from_values(ctx, src, 3, False, 4)
        src[:] = 0x1234567
        ref = O.batch_from_values(list(orig), 3, 4, want_leaves=False, want_digests=False)
        assert (b.cap == ref["cap"]).all()
        b.free()


def test_prove_rejects_inconsistent_fri_parameters(ctx, m):
    """p2b_prove validates the FRI parameters before deriving any length from them (no unsigned underflow, no
    misleading 'buffer too small')"""
    circ, digest, pis = make_case(5, ALL_GATES[:5], [(0, 4), (4, 5)], 63)
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 2, keep_values=True)
    for bad in (m.FriParams(3, 2, 6, 4, [4, 4, 4]),      # reduces below the blow-up
                m.FriParams(3, 7, 6, 4, [3, 2]),         # cap higher than a layer
                m.FriParams(3, 2, 6, 4, [3, 7])):        # arity out of range
        with pytest.raises(m.P2BError):
            m.prove_native(ctx, cd, cs, digest, circ.wire_values(), pis, bad, raw=True)
    ok = m.prove_native(ctx, cd, cs, digest, circ.wire_values(), pis, m.FriParams(3, 2, 6, 4, [3, 2]), raw=True)
    assert ok.size > 0
    cs.free()
    cd.free()


def test_gpu_openings_match_direct_evaluation(ctx, m):
    rng = np.random.default_rng(5)
    cols = [rng.integers(0, R.P, 64, dtype=np.uint64) for _ in range(5)]
    b = m.PolynomialBatch.from_coeffs(ctx, cols, 1, False, 0)
    z = [12345678901234567, 0xFFFFFFFF00000000]
    got = b.eval_ext(z, 1, 3)
    for i in range(3):
        e = R.horner_ext(cols[1 + i], R.Ext(*z))
        assert [int(got[i][0]), int(got[i][1])] == [e.a, e.b]
    with pytest.raises(m.P2BError):
        b.eval_ext(z, 3, 3)
    b.free()


def test_noncanonical_words_at_the_api(ctx, m):
    """GoldilocksField is a transparent u64 and the reference hands over non-canonical words in places
    (city_crypto/src/hash/qhashout.rs:149-152): coefficients >= p given to from_coeffs must evaluate, extend and fold
    like their canonical representatives (the constraint / prover kernels use strict canonical-in arithmetic, so every
    raw load has to be canonicalised first)."""
    P = R.P
    rng = np.random.default_rng(11)
    canonical = [rng.integers(0, P, 64, dtype=np.uint64) for _ in range(3)]
    raw = [c.copy() for c in canonical]
    for c in raw:  # sprinkle representatives x + p (possible for x < 2^32 - 1) and the extremes p, 2^64 - 1
        c[3] = np.uint64(7)
        c[9] = np.uint64(0)
        c[17] = np.uint64(2**32 - 2)
    canonical = [c.copy() for c in raw]
    for c in raw:
        c[3] = np.uint64(7 + P)
        c[9] = np.uint64(P)
        c[17] = np.uint64(2**64 - 1)
    a = m.PolynomialBatch.from_coeffs(ctx, canonical, 2, False, 1)
    b = m.PolynomialBatch.from_coeffs(ctx, raw, 2, False, 1)
    assert (a.cap == b.cap).all() and (a.leaves() == b.leaves()).all()
    z = [0xFFFFFFFF00000000, 3]
    assert (a.eval_ext(z) == b.eval_ext(z)).all()
    for i in range(3):
        e = R.horner_ext(canonical[i], R.Ext(*z))
        assert [int(b.eval_ext(z)[i][0]), int(b.eval_ext(z)[i][1])] == [e.a, e.b]
    a.free()
    b.free()
    # FRI commit phase on raw extension coefficients / values
    n, rate_bits = 64, 1
    N = n << rate_bits
    coeffs = np.zeros((N, 2), np.uint64)
    coeffs[:n] = np.stack([canonical[0], canonical[1]], axis=1)
    values = O.ext_coset_fft(coeffs, 7)
    coeffs_raw, values_raw = coeffs.copy(), values.copy()
    coeffs_raw[:n] = np.stack([raw[0], raw[1]], axis=1)
    small = values_raw < np.uint64(2**32 - 1)
    values_raw[small] += np.uint64(P)
    outs = []
    for cf, vl in ((coeffs, values), (coeffs_raw, values_raw)):
        ch = m.Challenger(ctx)
        ch.observe_elements([1, 2, 3])
        trees, final = m.fri_committed_trees(ctx, cf, vl, ch, [2, 1], rate_bits, 1)
        outs.append(([t.cap.copy() for t in trees], final.copy(), ch.export_state()[:12].tolist()))
    assert all((x == y).all() for x, y in zip(outs[0][0], outs[1][0]))
    assert (outs[0][1] == outs[1][1]).all() and outs[0][2] == outs[1][2]


def test_gpu_proof_2p14_rows_verifies(ctx, m):
    """a larger circuit (2^14 rows: multi-pass NTT path, three FRI layers): too slow to rebuild through the Python
    oracle prover, so the GPU proof goes straight through the restated verifier (which only needs the proof)"""
    fp = dict(rate_bits=3, cap_height=4, proof_of_work_bits=12, num_query_rounds=10, reduction_arity_bits=[4, 4, 4])
    gates = [(R.GATE_PUBLIC_INPUT, 0, 0), (R.GATE_NOOP, 0, 0), (R.GATE_CONSTANT, 2, 0), (R.GATE_ARITHMETIC, 20, 0)]
    pis = [1, 2, 3, 4, 5]
    circ = R.SyntheticCircuit(14, gates, [(0, 4)], 61, pi_hash=O.hash_no_pad(pis), link_prob=0.05)
    digest = [9, 8, 7, 6]
    cd = m.CircuitData(ctx, circ.desc())
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
    proof = m.prove_native(ctx, cd, cs, digest, circ.wire_values(), pis, m.FriParams(3, 4, 12, 10, [4, 4, 4]))
    assert V.verify(circ, cs.cap, digest, proof, fp)
    bad = dict(proof, openings=dict(proof["openings"]))
    w = bad["openings"]["wires"].copy()
    w[7][1] ^= np.uint64(1)
    bad["openings"]["wires"] = w
    with pytest.raises(V.VerificationError):
        V.verify(circ, cs.cap, digest, bad, fp)
    cs.free()
    cd.free()


def test_qbench_replay_native_job_loop(ctx, m, tmp_path):
    """tools/qbench_replay.cpp: the job DAG READ FROM the dumped block (tests/golden/example_dag.bin = the counters, goals
    and next-job lists of qbench_data/example.bin: 43 plonky2 jobs / 67 proofs, 3 Groth16 jobs skipped, 13 AggregateJobs
    joins) through the C++ mirror, (a) on worker threads with blocking p2b_prove and (b) with ONE host thread driving four
    contexts through p2b_prove_submit / collect.  Every proof must equal the words the CPU ORACLE computes for the case,
    every job must leave its bincode proof in the store, and the benchmark file must have the reference's
    QWorkerJobBenchmark format (city_rollup_common/src/qworker/job_id.rs:194-202)."""
    import json
    import shutil
    import subprocess
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import prove_bench as PB

    if not shutil.which("g++"):
        pytest.skip("no g++")
    circ, digest, pis = PB.build_case(10)
    expected, _ = V.oracle_prove_c(circ, digest, pis, FP_CITY)
    np.save(tmp_path / "expected.npy", expected)
    case = tmp_path / "case.bin"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_prove_case.py"), str(case), "10", str(tmp_path / "expected.npy")])
    exe = tmp_path / "qbench_replay"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", ROOT, os.path.join(ROOT, "tools", "qbench_replay.cpp"), "-L",
                           os.path.dirname(m.SO_PATH), "-lp2b", "-lpthread", "-o", str(exe)])
    dag = os.path.join(ROOT, "tests", "golden", "example_dag.bin")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(m.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    for mode in (["--contexts", "3"], ["--async", "4"]):
        out = tmp_path / "bench.json"
        res = subprocess.run([str(exe), "-i", str(case), "-d", dag, "-o", str(out), "-n", "2"] + mode, env=env,
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr + res.stdout
        summary = json.loads(res.stdout.strip().splitlines()[-1])
        assert summary["proving_jobs"] == summary["jobs_recorded"] == summary["stored_proofs"] == 2 * 43
        assert summary["jobs"] == 2 * 60  # + 3 Groth16 (skipped) + 13 AggregateJobs + NotifyOrchestratorComplete per block
        assert summary["proofs"] == 2 * 67 and summary["mismatching_proofs"] == 0
        if mode[0] == "--async":
            assert summary["host_threads"] == 1 and summary["contexts_per_gpu"] == 4
        bench = json.load(open(out))
        assert len(bench) == 2 * 43
        ids = {b["job_id"] for b in bench}
        assert len(ids) == 2 * 43 and all(len(i) == 48 and i.startswith("00") for i in ids)
        assert all(isinstance(b["duration"], int) for b in bench)


@pytest.mark.parametrize("env", [{"P2B_TREE_FUSE_LOG": "15"}, {"P2B_TREE_FUSE_LOG": "1"}, {"P2B_QUOT_EXT": "0"},
                                 {"P2B_GRAPH": "0"}, {"P2B_HASH_BLOCK": "128", "P2B_POW_BLOCKS": "3"}])
def test_tuning_knobs_do_not_change_results(env):
    """the environment knobs of INTEGRATION.md select other launch structures (Merkle levels fused from 2^15 digests /
    not at all, every gate at every point, no prove plans, other CTA sizes): smoke() — a commit and a complete proof,
    both against the oracle, the proof through the restated verifier — must pass under each of them"""
    import subprocess
    import sys

    res = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT,
                         env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "smoke ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
