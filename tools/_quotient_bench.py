"""development helper: time compute_quotient_polys (p2b_quotient_commit) on random batches of a given size
usage: _quotient_bench.py [log_rows=16] [gate_set=recursion|city] [reps=5]"""
import sys
import time

sys.path[:0] = ['.', 'tests', 'tools']
import numpy as np
import torch

import city_rollup_b200 as m
import plonk_ref as R

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gate_set = sys.argv[2] if len(sys.argv) > 2 else "recursion"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
gates, groups = (R.CITY_GATES, R.CITY_GROUPS) if gate_set == "city" else (R.RECURSION_GATES, R.RECURSION_GROUPS)
small = R.SyntheticCircuit(5, gates, groups, 3)
desc = dict(small.desc(), degree_bits=log_n)
c = m.Context(0)
cd = m.CircuitData(c, desc)
n = 1 << log_n
g = torch.Generator(device="cuda")
g.manual_seed(1)


def rand_batch(n_cols, keep):
    t = torch.empty((n_cols, n), dtype=torch.int64, device="cuda")
    t.random_(0, 2**62, generator=g)
    torch.cuda.synchronize()
    h = m.PolynomialBatch.from_values_device(c, t.data_ptr(), n_cols, log_n, 3, 4)
    return h


cs = rand_batch(desc["num_constants"] + desc["num_routed_wires"], False)
wires = rand_batch(desc["num_wires"], False)
zs = rand_batch(desc["num_challenges"] * (1 + desc["num_partial_products"]), False)
betas, gammas, alphas = [3, 5], [7, 11], [13, 17]
for i in range(reps + 1):
    c.profile_enable(True)
    c.profile_read()
    c.timer_start()
    q = m.compute_quotient_polys(c, cd, cs, [1, 2, 3, 4], wires, zs, betas, gammas, alphas, 3, 4)
    ms = c.timer_stop_ms()
    st, cnt = c.profile_read()
    c.profile_enable(False)
    q.free()
    if i:
        print("quotient+commit ms %.3f  constraint kernels ('other' stage) ms %.3f  launches %s" % (ms, st["other"], cnt["other"]))
