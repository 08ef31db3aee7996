import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import p2oracle as O
import plonk_ref as R
import city_rollup_b200 as m
from test_plonk_oracle import ALL_GATES
P = R.P
ctx = m.Context(0)
for name, gates, groups in [("noop", [ALL_GATES[1]], [(0, 1)]), ("const", [ALL_GATES[2]], [(0, 1)]), ("pi+noop", ALL_GATES[:2], [(0, 2)]),
                            ("arith", [ALL_GATES[3]], [(0, 1)]), ("first4", ALL_GATES[:4], [(0, 4)])]:
    circ = R.SyntheticCircuit(3, gates, groups, 21)
    rng = random.Random(5)
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(2)] for _ in range(3))
    d = circ.desc()
    cd = m.CircuitData(ctx, d)
    cs = m.PolynomialBatch.from_values(ctx, circ.constants_sigmas_values(), 3, False, 0, keep_values=True)
    wi = m.PolynomialBatch.from_values(ctx, circ.wire_values(), 3, False, 0, keep_values=True)
    zs = m.all_wires_permutation_partial_products(ctx, cd, cs, wi, betas, gammas, 3, 0)
    qt = m.compute_quotient_polys(ctx, cd, cs, circ.pi_hash, wi, zs, betas, gammas, alphas, 3, 0)
    o_cs = O.batch_from_values(circ.constants_sigmas_values(), 3, 0, want_digests=False)
    o_wi = O.batch_from_values(circ.wire_values(), 3, 0, want_digests=False)
    ref_zs = O.partial_products_and_zs(d, np.array(circ.wires, dtype=np.uint64), np.array(circ.sigmas, dtype=np.uint64), betas, gammas)
    o_zs = O.batch_from_values(list(ref_zs), 3, 0, want_digests=False)
    ref = O.compute_quotient_polys(d, 3, o_cs["leaves"], o_wi["leaves"], o_zs["leaves"], circ.pi_hash, betas, gammas, alphas)
    got = np.stack([qt.coeffs(c) for c in range(qt.n_cols)])
    n = circ.n
    for ch in range(2):
        rv = O.coset_fft(ref[ch * 8:(ch + 1) * 8].reshape(-1), 7)
        gv = O.coset_fft(got[ch * 8:(ch + 1) * 8].reshape(-1), 7)
        bad = np.nonzero(rv != gv)[0]
        print(name, "ch", ch, "coeff equal:", bool((ref[ch*8:(ch+1)*8] == got[ch*8:(ch+1)*8]).all()), "value mismatches:", len(bad), bad[:16].tolist())
