"""two complete synthetic proofs at the City shape (warm-up + one), for ncu launch lists"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from prove_bench import build_case, m
degree_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 12
circ, digest, pis = build_case(degree_bits)
c = m.Context(0)
cd = m.CircuitData(c, circ.desc())
cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), 3, False, 4, keep_values=True)
params = m.FriParams(3, 4, 16, 28, [4, 4] if degree_bits < 14 else [4, 4, 4])
wv = circ.wire_values()
m.prove_native(c, cd, cs, digest, wv, pis, params, raw=True)
l0 = c.launch_count()
m.prove_native(c, cd, cs, digest, wv, pis, params, raw=True)
print("launches", c.launch_count() - l0)
