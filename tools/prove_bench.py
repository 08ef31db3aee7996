#!/usr/bin/env python3
"""Throughput of complete synthetic proofs at the City Rollup shape (2^12 rows x 135 wires, rate 8, cap 4,
16-bit PoW, 28 queries, arities [4,4]): one context, then several contexts on separate host threads / CUDA
streams of the same GPU (the worker model of SURVEY.md §8(e): independent jobs, no collective).
Prints one JSON object.  Development / measurement tool (bench.py embeds the same measurement)."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import city_rollup_b200 as m  # noqa: E402


def build_case(degree_bits=12, seed=7, device=0, gate_set="city"):
    """synthetic satisfiable circuit over the whole gate set (tests/plonk_ref.py is the witness generator; the
    public-inputs hash comes from the GPU library, not from the oracle)"""
    import plonk_ref as R

    # The gate set a City Rollup op circuit carries (city_common_circuit/src/builder/pad_circuit.rs:31-55
    # add_city_common_gates + the in-tree u32 gates its gadgets add, city_common_circuit/src/u32/gates/*.rs): all 21
    # gate kinds in the six selector groups plonky2 forms for them (tests/test_plonk_oracle.py CITY_GATES); `recursion`
    # = the 13 gate types of the proofs stored in qbench_data/example.bin (135 wires, 123 gate constraints).
    from plonk_ref import CITY_GATES, CITY_GROUPS, RECURSION_GATES, RECURSION_GROUPS
    gates, groups = (CITY_GATES, CITY_GROUPS) if gate_set == "city" else (RECURSION_GATES, RECURSION_GROUPS)
    pis = [seed, 2, 3, 4]
    c = m.Context(device)
    pih = [int(x) for x in c.hash_no_pad(pis)]
    c.close()
    circ = R.SyntheticCircuit(degree_bits, gates, groups, seed, pi_hash=pih)
    return circ, [1, 2, 3, 4], pis


def run(n_ctx, n_proofs, circ, digest, pis, device=0, blocking=None, stats=None):
    """blocking: None = the library default (spin, or P2B_SYNC); stats (dict) receives the host CPU seconds per proof"""
    params = m.FriParams(3, 4, 16, 28, [4, 4])
    ctxs = [m.Context(device) for _ in range(n_ctx)]
    if blocking is not None:
        for c in ctxs:
            c.set_blocking_sync(blocking)
    state = []
    for c in ctxs:
        cd = m.CircuitData(c, circ.desc())
        cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
        state.append((cd, cs))
    wv = circ.wire_values()
    wires = []
    for c in ctxs:  # witness columns in pinned host memory, one matrix per context: a single DMA per proof
        w = c.pinned_empty((len(wv), wv[0].size))
        for j, col in enumerate(wv):
            w[j] = col
        wires.append(w)  # the (n_wires, n) matrix itself: its rows are the witness columns

    def worker(i, k):
        c = ctxs[i]
        cd, cs = state[i]
        for _ in range(k):
            m.prove_native(c, cd, cs, digest, wires[i], pis, params, raw=True)

    worker(0, 1)  # warm-up (tables, pools)
    for i in range(1, n_ctx):
        worker(i, 1)
    per = n_proofs // n_ctx
    th = [threading.Thread(target=worker, args=(i, per)) for i in range(n_ctx)]
    t0 = time.perf_counter()
    c0 = time.process_time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    if stats is not None:
        stats["cpu_s_per_proof"] = (time.process_time() - c0) / (per * n_ctx)
    for (cd, cs), c in zip(state, ctxs):
        cs.free()
        cd.free()
        c.close()
    return per * n_ctx / dt, dt / per * 1e3


if __name__ == "__main__":
    circ, digest, pis = build_case()
    out = {}
    for n_ctx in (1, 2, 4, 8, 12):
        for blocking in (False, True):
            st = {}
            pps, ms = run(n_ctx, 48 * n_ctx, circ, digest, pis, blocking=blocking, stats=st)
            out[f"ctx{n_ctx}_{'block' if blocking else 'spin'}"] = {
                "proofs_per_s": round(pps, 2), "ms_per_proof_per_ctx": round(ms, 2),
                "host_cpu_ms_per_proof": round(st["cpu_s_per_proof"] * 1e3, 2)}
    # stage profile of one proof
    c = m.Context(0)
    cd = m.CircuitData(c, circ.desc())
    cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
    params = m.FriParams(3, 4, 16, 28, [4, 4])
    m.prove_native(c, cd, cs, digest, circ.wire_values(), pis, params)
    c.profile_enable(True)
    c.profile_read()
    l0 = c.launch_count()
    wv = circ.wire_values()
    t0 = time.perf_counter()
    m.prove_native(c, cd, cs, digest, wv, pis, params, raw=True)
    out["one_proof_wall_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    ms, cnt = c.profile_read()
    out["stage_ms"] = {k: round(v, 3) for k, v in ms.items()}
    out["launches_per_proof"] = c.launch_count() - l0
    print(json.dumps(out))
