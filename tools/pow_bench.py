#!/usr/bin/env python3
"""Latency of p2b_fri_pow (minimal witness) over a handful of transcripts.  Development tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import city_rollup_b200 as m  # noqa: E402

ctx = m.Context(0)
for bits in (0, 8, 12, 16, 18, 20):
    ts, ws = [], []
    for seed in range(16):
        ch = m.Challenger(ctx)
        ch.observe_elements([seed, 6, 7, 8, 9])
        ctx.synchronize()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        w = m.fri_proof_of_work(ctx, ch, bits)
        ts.append((time.perf_counter() - t0) * 1e3)
        ws.append(w)
        nl = ctx.launch_count() - l0
        ch.free()
    print("bits", bits, "launches/call", nl, "mean ms %.3f" % (sum(ts[4:]) / len(ts[4:])), "min %.3f max %.3f" % (min(ts[4:]), max(ts[4:])),
          "witnesses", ws[:6])
