#!/usr/bin/env python3
"""Writes the case file tools/prove_bench.cpp reads: the synthetic City-shaped circuit of tools/prove_bench.py
(description, constants|sigmas values, witness columns, FRI parameters) and the proof words p2b_prove returns
for it (computed here on the GPU, or — argv[3], what tests/test_gpu_prove.py does — taken from a .npy the caller
produced with the CPU oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from prove_bench import build_case, m  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prove_case.bin")
degree_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 12
circ, digest, pis = build_case(degree_bits)
d = circ.desc()
params = m.FriParams(3, 4, 16, 28, [4, 4])
c = m.Context(0)
cd = m.CircuitData(c, d)
cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
words = m.prove_native(c, cd, cs, digest, circ.wire_values(), pis, params, raw=True)
# tests hand over the expected words computed by the CPU oracle instead (argv[3] = .npy): the replay then checks every
# GPU proof against an oracle-built proof, not against another GPU proof
if len(sys.argv) > 3:
    words = np.load(sys.argv[3]).astype(np.uint64)
u = lambda xs: np.array([int(x) for x in xs], dtype=np.uint64)
hdr = [0x70326263617365, d["degree_bits"], d["num_wires"], d["num_routed_wires"], d["num_constants"], d["num_selectors"],
       d["num_challenges"], d["quotient_degree_factor"], d["num_partial_products"], d["num_gate_constraints"],
       len(d["gates"]), len(pis), len(words), 0, 0, 0]
with open(out, "wb") as f:
    f.write(u(hdr).tobytes())
    for g in d["gates"]:
        f.write(u([g["kind"], g["p0"], g["p1"], g["selector_index"], g["group_start"], g["group_end"], g["row"]]).tobytes())
    f.write(u(d["k_is"]).tobytes())
    for col in circ.constants_sigmas_values() + circ.wire_values():
        f.write(np.ascontiguousarray(col, dtype=np.uint64).tobytes())
    f.write(u(digest).tobytes())
    f.write(u(pis).tobytes())
    f.write(u([params.rate_bits, params.cap_height, params.proof_of_work_bits, params.num_query_rounds,
               len(params.reduction_arity_bits)] + params.reduction_arity_bits + [0] * (16 - len(params.reduction_arity_bits))).tobytes())
    f.write(np.ascontiguousarray(words).tobytes())
print("wrote", out, os.path.getsize(out), "bytes;", len(words), "proof words")
