"""Test-side restatement of plonky2 0.2.2's proof flow for circuits without lookups / blinding:

  * oracle_prove: plonk/prover.rs::prove_with_partition_witness from the filled witness onwards, built ONLY from
    oracle primitives (oracle/*.c through p2oracle.py) and exact Python integer arithmetic — the word-for-word
    reference the CUDA prover is compared with;
  * verify: plonk/verifier.rs::verify_with_challenges + fri/verifier.rs::verify_fri_proof (get_challenges,
    eval_vanishing_poly, fri_verify_initial_proof, fri_combine_initial, compute_evaluation, final_poly check,
    proof-of-work check), an independent consumer of a proof.

The FRI conventions used here (leaf order, fold, the two opening batches and their alpha powers) are the ones
pinned on the ten proofs stored in the reference's qbench_data/example.bin (tests/test_oracle_golden.py).
"""
import numpy as np

import p2oracle as O
import plonk_ref as R

P = R.P
Ext = R.Ext


def bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def _g(degree_bits):
    return pow(pow(7, (P - 1) >> 32, P), 1 << (32 - degree_bits), P) if degree_bits else 1


def ext_list(a):
    return [Ext(int(x[0]), int(x[1])) for x in a]


def ext_arr(v):
    return np.array([[e.a, e.b] for e in v], dtype=np.uint64).reshape(-1, 2)


def reduce_with_powers(vals, alpha):
    acc = Ext(0)
    for v in reversed(vals):
        acc = acc * alpha + v
    return acc


def eval_polys_ext(coeff_rows, z):
    return [R.horner_ext(c, z) for c in coeff_rows]


OPENING_ORDER = ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys")


# ------------------------------------------------------------------------------------------------ oracle prover
def oracle_prove(circ, circuit_digest, public_inputs, fp):
    """-> (proof dict shaped like city_rollup_b200.prove's, constants_sigmas cap)"""
    d = circ.desc()
    nch, rb, cap_h = circ.num_challenges, fp["rate_bits"], fp["cap_height"]
    n, log_n = circ.n, circ.degree_bits
    pih = O.hash_no_pad(public_inputs)
    assert [int(x) for x in pih] == [int(x) for x in circ.pi_hash], "circuit must be built with pi_hash = hash(public_inputs)"
    cs = O.batch_from_values(circ.constants_sigmas_values(), rb, cap_h)
    wi = O.batch_from_values(circ.wire_values(), rb, cap_h)
    ch = O.Challenger()
    ch.observe(circuit_digest)
    ch.observe(pih)
    ch.observe(wi["cap"])
    betas, gammas = ch.get_n(nch), ch.get_n(nch)
    zs_vals = O.partial_products_and_zs(d, np.array(circ.wires, dtype=np.uint64), np.array(circ.sigmas, dtype=np.uint64),
                                        betas, gammas)
    zs = O.batch_from_values(list(zs_vals), rb, cap_h)
    ch.observe(zs["cap"])
    alphas = ch.get_n(nch)
    chunks = O.compute_quotient_polys(d, rb, cs["leaves"], wi["leaves"], zs["leaves"], pih, betas, gammas, alphas)
    qt = O.batch_from_coeffs(list(chunks), rb, cap_h)
    qt["coeffs"] = chunks
    ch.observe(qt["cap"])
    zeta_l = ch.get_ext()
    zeta = Ext(*zeta_l)
    zeta_next = zeta * _g(log_n)
    nc = circ.num_constants
    cs_z = eval_polys_ext(cs["coeffs"], zeta)
    zs_z = eval_polys_ext(zs["coeffs"], zeta)
    openings = dict(constants=cs_z[:nc], plonk_sigmas=cs_z[nc:], wires=eval_polys_ext(wi["coeffs"], zeta),
                    plonk_zs=zs_z[:nch], partial_products=zs_z[nch:],
                    quotient_polys=eval_polys_ext(chunks, zeta),
                    plonk_zs_next=eval_polys_ext(zs["coeffs"][:nch], zeta_next))
    for k in OPENING_ORDER + ("plonk_zs_next",):
        ch.observe(ext_arr(openings[k]))
    # ---- PolynomialBatch::prove_openings
    alpha = Ext(*ch.get_ext())
    oracles = [cs, wi, zs, qt]
    batches = [(zeta, [c for o in oracles for c in o["coeffs"]]), (zeta_next, list(zs["coeffs"][:nch]))]
    final = [Ext(0)] * n
    for point, polys in batches:
        comp = [Ext(0)] * n
        apow = Ext(1)
        for poly in polys:  # alpha.reduce_polys_base
            comp = [c + apow * int(x) for c, x in zip(comp, poly)]
            apow = apow * alpha
        bs, acc = [], Ext(0)  # divide_by_linear
        for c in reversed(comp):
            acc = acc * point + c
            bs.append(acc)
        bs.pop()
        bs.reverse()
        quot = bs + [Ext(0)]
        final = [f * apow + q for f, q in zip(final, quot)]  # shift_poly by alpha^len(polys), then add
    N = n << rb
    coeffs = np.zeros((N, 2), np.uint64)
    coeffs[:n] = ext_arr(final)
    values = O.ext_coset_fft(coeffs, 7)
    fri = O.fri_committed_trees(coeffs, values, fp["reduction_arity_bits"], ch, rb, cap_h)
    pow_witness = O.fri_proof_of_work(ch, fp["proof_of_work_bits"])
    log_N = log_n + rb
    rounds = []
    for _ in range(fp["num_query_rounds"]):
        x = ch.get() % N
        initial = [(o["leaves"][x].copy(), O.merkle_prove(o["digests"], N, cap_h, x)) for o in oracles]
        steps = []
        ln = N
        for a, leaves, dg in zip(fp["reduction_arity_bits"], fri["leaves"], fri["digests"]):
            ln >>= a
            x >>= a
            steps.append((leaves[x].reshape(-1, 2).copy(), O.merkle_prove(dg, ln, cap_h, x)))
        rounds.append(dict(initial_trees_proof=initial, steps=steps))
    proof = dict(wires_cap=wi["cap"], plonk_zs_partial_products_cap=zs["cap"], quotient_polys_cap=qt["cap"],
                 openings={k: ext_arr(v) for k, v in openings.items()},
                 opening_proof=dict(commit_phase_merkle_caps=list(fri["caps"]), query_round_proofs=rounds,
                                    final_poly=fri["final_poly"], pow_witness=pow_witness),
                 public_inputs=[int(x) for x in public_inputs])
    return proof, cs["cap"]


# ------------------------------------------------------------------------------------------------ verifier
class VerificationError(AssertionError):
    pass


def _ensure(cond, msg):
    if not cond:
        raise VerificationError(msg)


def verify(circ, cs_cap, circuit_digest, proof, fp):
    nch, rb, cap_h = circ.num_challenges, fp["rate_bits"], fp["cap_height"]
    log_n, n = circ.degree_bits, circ.n
    log_N, N = log_n + rb, n << rb
    op = {k: ext_list(v) for k, v in proof["openings"].items()}
    fri = proof["opening_proof"]
    # ---- get_challenges
    ch = O.Challenger()
    pih = O.hash_no_pad(proof["public_inputs"])
    ch.observe(circuit_digest)
    ch.observe(pih)
    ch.observe(proof["wires_cap"])
    betas, gammas = ch.get_n(nch), ch.get_n(nch)
    ch.observe(proof["plonk_zs_partial_products_cap"])
    alphas = ch.get_n(nch)
    ch.observe(proof["quotient_polys_cap"])
    zeta = Ext(*ch.get_ext())
    for k in OPENING_ORDER + ("plonk_zs_next",):
        ch.observe(np.asarray(proof["openings"][k], dtype=np.uint64))
    fri_alpha = Ext(*ch.get_ext())
    fri_betas = []
    for cap in fri["commit_phase_merkle_caps"]:
        ch.observe(cap)
        fri_betas.append(ch.get_ext())
    ch.observe(np.asarray(fri["final_poly"], dtype=np.uint64))
    ch.observe([fri["pow_witness"]])
    pow_response = ch.get()
    _ensure(pow_response >> (64 - fp["proof_of_work_bits"]) == 0 if fp["proof_of_work_bits"] else True, "proof of work")
    indices = [ch.get() % N for _ in range(fp["num_query_rounds"])]
    # ---- PLONK: vanishing(zeta) = Z_H(zeta) * sum_i zeta^(n i) t_i(zeta)
    circ_pi = list(circ.pi_hash)
    _ensure([int(x) for x in pih] == [int(x) for x in circ_pi], "public inputs hash")
    npp = circ.num_pp
    pps = [op["partial_products"][i * npp:(i + 1) * npp] for i in range(nch)]
    van, z_h, zeta_n = R.eval_vanishing_poly_ext(circ, zeta, op["constants"], op["plonk_sigmas"], op["wires"], op["plonk_zs"],
                                                 op["plonk_zs_next"], pps, betas, gammas, alphas)
    for i in range(nch):
        t = reduce_with_powers(op["quotient_polys"][i * circ.qdf:(i + 1) * circ.qdf], zeta_n)
        _ensure(van[i] == z_h * t, f"vanishing polynomial identity, challenge {i}")
    # ---- FRI
    n_final = (N >> sum(fp["reduction_arity_bits"])) >> rb
    _ensure(len(fri["final_poly"]) == n_final, "final polynomial length")
    zeta_next = zeta * _g(log_n)
    batch0 = [v for k in OPENING_ORDER for v in op[k]]
    batch1 = op["plonk_zs_next"]
    red0, red1 = reduce_with_powers(batch0, fri_alpha), reduce_with_powers(batch1, fri_alpha)
    caps = [cs_cap, proof["wires_cap"], proof["plonk_zs_partial_products_cap"], proof["quotient_polys_cap"]]
    w = O.root_of_unity(log_N)
    final_poly = ext_list(fri["final_poly"])
    _ensure(len(fri["query_round_proofs"]) == fp["num_query_rounds"], "number of query rounds")
    for x_index, rnd in zip(indices, fri["query_round_proofs"]):
        init = rnd["initial_trees_proof"]
        for (leaf, sib), cap in zip(init, caps):
            _ensure(O.merkle_verify(leaf, x_index, sib, cap), "initial Merkle proof")
        x = 7 * pow(w, bitrev(x_index, log_N), P) % P
        ev0 = [Ext(int(v)) for leaf, _ in init for v in leaf]
        ev1 = [Ext(int(v)) for v in init[2][0][:nch]]
        _ensure(len(ev0) == len(batch0), "leaf widths")
        s = (reduce_with_powers(ev0, fri_alpha) - red0) * (Ext(x) - zeta).inv()
        s = s * fri_alpha ** len(ev1) + (reduce_with_powers(ev1, fri_alpha) - red1) * (Ext(x) - zeta_next).inv()
        old_eval = s
        xi = x_index
        for i, a in enumerate(fp["reduction_arity_bits"]):
            evals, sib = rnd["steps"][i]
            coset, within = xi >> a, xi & ((1 << a) - 1)
            _ensure(Ext(int(evals[within][0]), int(evals[within][1])) == old_eval, f"FRI consistency at layer {i}")
            e = O.fri_compute_evaluation(x, within, a, evals, fri_betas[i])
            old_eval = Ext(*e)
            _ensure(O.merkle_verify(np.asarray(evals, dtype=np.uint64).reshape(-1), coset, sib, fri["commit_phase_merkle_caps"][i]),
                    f"FRI layer {i} Merkle proof")
            x = pow(x, 1 << a, P)
            xi = coset
        _ensure(reduce_with_powers(final_poly, Ext(x)) == old_eval, "final polynomial evaluation")
    return True


def proofs_equal(a, b):
    """word-for-word comparison of two proof dicts; returns the name of the first differing field or None"""
    for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap"):
        if not (np.asarray(a[k]) == np.asarray(b[k])).all():
            return k
    for k in a["openings"]:
        if not (np.asarray(a["openings"][k], dtype=np.uint64) == np.asarray(b["openings"][k], dtype=np.uint64)).all():
            return "openings." + k
    fa, fb = a["opening_proof"], b["opening_proof"]
    for i, (x, y) in enumerate(zip(fa["commit_phase_merkle_caps"], fb["commit_phase_merkle_caps"])):
        if not (np.asarray(x) == np.asarray(y)).all():
            return f"commit_phase_merkle_caps[{i}]"
    if not (np.asarray(fa["final_poly"]) == np.asarray(fb["final_poly"])).all():
        return "final_poly"
    if fa["pow_witness"] != fb["pow_witness"]:
        return "pow_witness"
    for q, (ra, rb_) in enumerate(zip(fa["query_round_proofs"], fb["query_round_proofs"])):
        for o, ((la, sa), (lb, sb)) in enumerate(zip(ra["initial_trees_proof"], rb_["initial_trees_proof"])):
            if not ((np.asarray(la) == np.asarray(lb)).all() and (np.asarray(sa) == np.asarray(sb)).all()):
                return f"query {q} initial tree {o}"
        for l, ((ea, sa), (eb, sb)) in enumerate(zip(ra["steps"], rb_["steps"])):
            if not ((np.asarray(ea) == np.asarray(eb)).all() and (np.asarray(sa) == np.asarray(sb)).all()):
                return f"query {q} step {l}"
    if a["public_inputs"] != b["public_inputs"]:
        return "public_inputs"
    return None


def flatten_proof(proof):
    """proof dict -> the flat u64 words p2b_prove / p2o_prove write (include/p2b.h: ProofWithPublicInputs' field
    order, no length prefixes)"""
    parts = [np.asarray(proof[k], dtype=np.uint64).reshape(-1)
             for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap")]
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys"):
        parts.append(np.asarray(proof["openings"][k], dtype=np.uint64).reshape(-1))
    fri = proof["opening_proof"]
    for cap in fri["commit_phase_merkle_caps"]:
        parts.append(np.asarray(cap, dtype=np.uint64).reshape(-1))
    for rnd in fri["query_round_proofs"]:
        for leaf, sib in rnd["initial_trees_proof"]:
            parts += [np.asarray(leaf, dtype=np.uint64).reshape(-1), np.asarray(sib, dtype=np.uint64).reshape(-1)]
        for ev, sib in rnd["steps"]:
            parts += [np.asarray(ev, dtype=np.uint64).reshape(-1), np.asarray(sib, dtype=np.uint64).reshape(-1)]
    parts.append(np.asarray(fri["final_poly"], dtype=np.uint64).reshape(-1))
    parts.append(np.array([fri["pow_witness"]], dtype=np.uint64))
    parts.append(np.array(proof["public_inputs"], dtype=np.uint64).reshape(-1))
    return np.concatenate(parts)


def oracle_prove_c(circ, circuit_digest, public_inputs, fp):
    """the same proof from the C composition (oracle/prove.c::p2o_prove) -> (flat words, constants_sigmas cap)"""
    pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), fp)
    words = pd.prove(circuit_digest, circ.wire_values(), public_inputs)
    cap = pd.cap.copy()
    pd.free()
    return words, cap


def parse_proof(circ, fp, words, n_public_inputs):
    """flat words -> proof dict (inverse of flatten_proof)"""
    words = np.asarray(words, dtype=np.uint64)
    nch, nc, nr, nw = circ.num_challenges, circ.num_constants, circ.num_routed, circ.num_wires
    npp, qdf, cap_h, rb = circ.num_pp, circ.qdf, fp["cap_height"], fp["rate_bits"]
    log_N = circ.degree_bits + rb
    pos = 0

    def take(k, shape):
        nonlocal pos
        a = words[pos:pos + k].reshape(shape).copy()
        pos += k
        return a

    cw = 4 << cap_h
    proof = dict(wires_cap=take(cw, (-1, 4)), plonk_zs_partial_products_cap=take(cw, (-1, 4)), quotient_polys_cap=take(cw, (-1, 4)))
    op = {}
    for k, cnt in (("constants", nc), ("plonk_sigmas", nr), ("wires", nw), ("plonk_zs", nch), ("plonk_zs_next", nch),
                   ("partial_products", nch * npp), ("quotient_polys", nch * qdf)):
        op[k] = take(2 * cnt, (-1, 2))
    proof["openings"] = op
    caps = [take(cw, (-1, 4)) for _ in fp["reduction_arity_bits"]]
    rounds = []
    for _ in range(fp["num_query_rounds"]):
        initial = []
        for w in (nc + nr, nw, nch * (1 + npp), nch * qdf):
            leaf = take(w, (-1,))
            initial.append((leaf, take(4 * (log_N - cap_h), (-1, 4))))
        steps, lc = [], log_N
        for a in fp["reduction_arity_bits"]:
            lc -= a
            ev = take(2 << a, (-1, 2))
            steps.append((ev, take(4 * (lc - cap_h), (-1, 4))))
        rounds.append(dict(initial_trees_proof=initial, steps=steps))
    n_final = ((1 << log_N) >> sum(fp["reduction_arity_bits"])) >> rb
    final_poly = take(2 * n_final, (-1, 2))
    pow_witness = int(take(1, (-1,))[0])
    proof["opening_proof"] = dict(commit_phase_merkle_caps=caps, query_round_proofs=rounds, final_poly=final_poly,
                                  pow_witness=pow_witness)
    proof["public_inputs"] = [int(x) for x in take(n_public_inputs, (-1,))]
    assert pos == words.size
    return proof
