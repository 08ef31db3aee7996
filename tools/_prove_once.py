"""development helper: a few City-shape proofs on one context (for ncu launch lists / captures)"""
import sys
sys.path[:0] = ['.', 'tests', 'tools']
import numpy as np
import city_rollup_b200 as m
import prove_bench as PB

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
circ, digest, pis = PB.build_case()
c = m.Context(0)
cd = m.CircuitData(c, circ.desc())
cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
params = m.FriParams(3, 4, 16, 28, [4, 4])
wv = np.stack(circ.wire_values())
for i in range(n):
    l0 = c.launch_count()
    m.prove_native(c, cd, cs, digest, wv, pis, params, raw=True)
    print("launches per proof", c.launch_count() - l0, file=sys.stderr)
