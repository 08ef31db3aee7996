"""GPU: the complete prover flow behind the C ABI (wires commit -> Z / partial products -> quotient -> openings ->
prove_openings: final polynomial, FRI commit phase, proof of work, query rounds) against the oracle-built proof
(tests/verifier_ref.py::oracle_prove), word for word, and through the restated plonky2 verifier."""
import os

import numpy as np
import pytest

import p2oracle as O
import plonk_ref as R
import verifier_ref as V
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
    """tools/qbench_replay.cpp: the reference's level-counter job DAG (43 jobs / 67 proofs of a block shaped like
    qbench_data/example.bin) through the C++ mirror on worker threads; every proof must equal the expected words, every
    job must leave its bincode proof in the store, and the benchmark file must have the reference's
    QWorkerJobBenchmark format (city_rollup_common/src/qworker/job_id.rs:194-202)."""
    import json
    import shutil
    import subprocess
    import sys

    if not shutil.which("g++"):
        pytest.skip("no g++")
    case = tmp_path / "case.bin"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_prove_case.py"), str(case), "10"])
    exe = tmp_path / "qbench_replay"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", ROOT, os.path.join(ROOT, "tools", "qbench_replay.cpp"), "-L",
                           os.path.dirname(m.SO_PATH), "-lp2b", "-lpthread", "-o", str(exe)])
    out = tmp_path / "bench.json"
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(m.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res = subprocess.run([str(exe), "-i", str(case), "-o", str(out), "-n", "2", "--contexts", "3"], env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr + res.stdout
    summary = json.loads(res.stdout.strip().splitlines()[-1])
    assert summary["jobs"] == summary["jobs_recorded"] == summary["stored_proofs"] == 2 * 43
    assert summary["proofs"] == 2 * 67 and summary["mismatching_proofs"] == 0
    bench = json.load(open(out))
    assert len(bench) == 2 * 43
    ids = {b["job_id"] for b in bench}
    assert len(ids) == 2 * 43 and all(len(i) == 48 and i.startswith("00") for i in ids)
    assert all(isinstance(b["duration"], int) for b in bench)
