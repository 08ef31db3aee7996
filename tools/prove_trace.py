"""Phase-by-phase wall clock of one synthetic City-shape proof (P2B_TRACE=1; pinned witness).  Development tool."""
import os, sys
os.environ["P2B_TRACE"] = "1"
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
w = c.pinned_empty((len(wv), wv[0].size))
for j, col in enumerate(wv):
    w[j] = col
cols = [w[j] for j in range(len(wv))]
for i in range(3):
    sys.stderr.write("---- proof %d\n" % i)
    m.prove_native(c, cd, cs, digest, cols, pis, params, raw=True)
