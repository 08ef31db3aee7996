"""development helper: BASELINE.json configs[2] beside the 2^20 x 400 commit (tools/_commit_once.py 20 400): the quotient
stage at 2^20 rows (recursion gate set, random batches) and fri_committed_trees on a 2^23-point extension polynomial with
arities [4,4,4,4] (host coefficient / value planes: the upload is part of the call)."""
import subprocess
import sys
import time

sys.path[:0] = ['.', 'tests', 'tools']
import numpy as np

import city_rollup_b200 as m
from util import rand_felts

subprocess.run([sys.executable, "tools/_quotient_bench.py", "20", "recursion", "2"], check=False)
log_n, rate_bits, cap_height, arity_bits = 20, 3, 4, [4, 4, 4, 4]
n = 1 << log_n
N = n << rate_bits
ctx = m.Context(0)
coeffs = np.zeros((N, 2), np.uint64)
coeffs[:n] = rand_felts(2023, (n, 2))
values = rand_felts(2024, (N, 2))  # timing only: the fold works on the coefficients, the first layer's leaves on these
for rep in range(3):
    gc = m.Challenger(ctx)
    gc.observe_elements([1, 2, 3, 4])
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.timer_start()
    trees, final = m.fri_committed_trees(ctx, coeffs, values, gc, arity_bits, rate_bits, cap_height)
    ms = ctx.timer_stop_ms()
    wall = (time.perf_counter() - t0) * 1e3
    print("fri_committed_trees 2^23 points, arities [4,4,4,4]: %.3f ms on the device (incl. the upload of 2 x 134 MB), %.1f ms wall" % (ms, wall))
    for t in trees:
        t.free()
