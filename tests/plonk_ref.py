"""Test-side helpers for the PLONK stages: a synthetic circuit builder with a witness
generator, and an INDEPENDENT pure-Python restatement of plonky2 0.2.2's verifier-side `eval_vanishing_poly`
(plonk/vanishing_poly.rs) over the quadratic extension, written against the scalar `eval_unfiltered` of every
gate (upstream gates/*.rs; in-tree city_common_circuit/src/u32/gates/*.rs).  It is used to check the verifier
identity  sum_k alpha^k term_k(zeta) = Z_H(zeta) * sum_i zeta^(n i) t_i(zeta)  on the quotient chunks produced by
the oracle / the CUDA path, which pins both against a definition that shares no code with either.

No real City Rollup circuit can be built here (the Rust CircuitBuilder is not available), so circuits are
synthetic: random rows of the closed gate set with random copy constraints.
"""
import importlib.util
import os
import random

import numpy as np

P = 2**64 - 2**32 + 1
W = 7  # X^2 = 7

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location(
    "gen_poseidon_fast_tables", os.path.join(_HERE, "..", "tools", "gen_poseidon_fast_tables.py"))
PF = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(PF)  # derives FIRST / POST / INIT / W_HATS / VS from MDS + round constants, self-checks

(GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_POSEIDON, GATE_BASE_SUM, GATE_U32_ARITHMETIC,
 GATE_U32_ADD_MANY, GATE_U32_SUBTRACTION, GATE_U32_RANGE_CHECK, GATE_U32_INTERLEAVE, GATE_UNINTERLEAVE_TO_U32,
 GATE_UNINTERLEAVE_TO_B32, GATE_COMPARISON, GATE_ARITHMETIC_EXT, GATE_MUL_EXT, GATE_REDUCING, GATE_REDUCING_EXT,
 GATE_RANDOM_ACCESS, GATE_POSEIDON_MDS, GATE_COSET_INTERPOLATION) = range(21)
UNUSED_SELECTOR = 2**32 - 1


class Ext:
    """F_p[X]/(X^2 - 7)"""
    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a, self.b = a % P, b % P

    @staticmethod
    def of(x):
        return x if isinstance(x, Ext) else Ext(x)

    def __add__(self, o):
        o = Ext.of(o)
        return Ext(self.a + o.a, self.b + o.b)

    __radd__ = __add__

    def __sub__(self, o):
        o = Ext.of(o)
        return Ext(self.a - o.a, self.b - o.b)

    def __rsub__(self, o):
        return Ext.of(o) - self

    def __mul__(self, o):
        o = Ext.of(o)
        return Ext(self.a * o.a + W * self.b * o.b, self.a * o.b + self.b * o.a)

    __rmul__ = __mul__

    def inv(self):
        nrm = pow((self.a * self.a - W * self.b * self.b) % P, P - 2, P)
        return Ext(self.a * nrm, -self.b * nrm)

    def __pow__(self, e):
        r, b = Ext(1), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def __eq__(self, o):
        o = Ext.of(o)
        return self.a == o.a and self.b == o.b

    def __repr__(self):
        return f"Ext({self.a}, {self.b})"


class Fp:
    """base field with the same operator surface as Ext (used to check a witness row by row)"""
    __slots__ = ("a",)

    def __init__(self, a):
        self.a = a % P

    @staticmethod
    def of(x):
        return x if isinstance(x, Fp) else Fp(x)

    def __add__(self, o):
        return Fp(self.a + Fp.of(o).a)

    __radd__ = __add__

    def __sub__(self, o):
        return Fp(self.a - Fp.of(o).a)

    def __rsub__(self, o):
        return Fp.of(o) - self

    def __mul__(self, o):
        return Fp(self.a * Fp.of(o).a)

    __rmul__ = __mul__

    def __eq__(self, o):
        return self.a == Fp.of(o).a


class Alg:
    """ExtensionAlgebra element c0 + c1 Y, Y^2 = 7, over any field-like component type (plonky2
    field/extension/algebra.rs): how a D = 2 extension element stored in two wires is multiplied both by the
    prover (components in F) and by the verifier (components in F_ext)."""
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def __add__(self, o):
        return Alg(self.c0 + o.c0, self.c1 + o.c1)

    def __sub__(self, o):
        return Alg(self.c0 - o.c0, self.c1 - o.c1)

    def __mul__(self, o):
        if isinstance(o, Alg):
            return Alg(self.c0 * o.c0 + self.c1 * o.c1 * W, self.c0 * o.c1 + self.c1 * o.c0)
        return Alg(self.c0 * o, self.c1 * o)  # scalar_mul

    def parts(self):
        return [self.c0, self.c1]


def _ext_at(w, start):
    return Alg(w[start], w[start + 1])


# ------------------------------------------------------------------------------------------ gates (scalar form)
def _mds(s):
    out = []
    for r in range(12):
        acc = s[r] * 0
        for i in range(12):
            acc = acc + s[(i + r) % 12] * PF.CIRC[i]
        if r == 0:
            acc = acc + s[0] * 8
        out.append(acc)
    return out


def _sbox(x):
    x2 = x * x
    x4 = x2 * x2
    return x * x2 * x4


def eval_gate(kind, p0, p1, w, consts, pi_hash):
    """constraints of one gate at one point; w / consts / pi_hash are sequences of field-like values"""
    c = []
    if kind == GATE_NOOP:
        pass
    elif kind == GATE_CONSTANT:
        for i in range(p0):
            c.append(consts[i] - w[i])
    elif kind == GATE_PUBLIC_INPUT:
        for i in range(4):
            c.append(w[i] - pi_hash[i])
    elif kind == GATE_ARITHMETIC:
        for i in range(p0):
            c.append(w[4 * i + 3] - (w[4 * i] * w[4 * i + 1] * consts[0] + w[4 * i + 2] * consts[1]))
    elif kind == GATE_BASE_SUM:
        acc = w[0] * 0
        for i in reversed(range(p0)):
            acc = acc * 2 + w[1 + i]
        c.append(acc - w[0])
        for i in range(p0):
            c.append(w[1 + i] * (w[1 + i] - 1))
    elif kind == GATE_POSEIDON:
        SWAP, DELTA, FULL0, PARTIAL, FULL1 = 24, 25, 29, 65, 87
        swap = w[SWAP]
        c.append(swap * (swap - 1))
        for i in range(4):
            c.append(swap * (w[i + 4] - w[i]) - w[DELTA + i])
        s = [None] * 12
        for i in range(4):
            s[i] = w[i] + w[DELTA + i]
            s[i + 4] = w[i + 4] - w[DELTA + i]
        for i in range(8, 12):
            s[i] = w[i]
        rnd = 0
        for r in range(4):
            s = [s[i] + PF.RC[12 * rnd + i] for i in range(12)]
            if r != 0:
                for i in range(12):
                    sin = w[FULL0 + 12 * (r - 1) + i]
                    c.append(s[i] - sin)
                    s[i] = sin
            s = _mds([_sbox(x) for x in s])
            rnd += 1
        s = [s[i] + PF.first_const[i] for i in range(12)]
        t = [s[0]]
        for i in range(11):
            acc = s[0] * 0
            for j in range(11):
                acc = acc + s[1 + j] * PF.INIT[i][j]
            t.append(acc)
        s = t
        for r in range(22):
            sin = w[PARTIAL + r]
            c.append(s[0] - sin)
            s[0] = _sbox(sin)
            if r < 21:
                s[0] = s[0] + PF.post[r]
            d = s[0] * PF.M[0][0]
            for i in range(1, 12):
                d = d + s[i] * PF.W_HATS[r][i - 1]
            s = [d] + [s[i] + s[0] * PF.VS[r][i - 1] for i in range(1, 12)]
        rnd += 22
        for r in range(4):
            s = [s[i] + PF.RC[12 * rnd + i] for i in range(12)]
            for i in range(12):
                sin = w[FULL1 + 12 * r + i]
                c.append(s[i] - sin)
                s[i] = sin
            s = _mds([_sbox(x) for x in s])
            rnd += 1
        for i in range(12):
            c.append(s[i] - w[12 + i])
    elif kind == GATE_U32_ARITHMETIC:  # arithmetic_u32.rs:88-150
        ops = p0
        for i in range(ops):
            m0, m1, ad, lo, hi, inv = (w[6 * i + k] for k in range(6))
            computed = m0 * m1 + ad
            c.append((inv * (0xFFFFFFFF - hi) - 1) * lo)
            c.append(hi * (1 << 32) + lo - computed)
            clo, chi = w[0] * 0, w[0] * 0
            for j in reversed(range(32)):
                limb = w[6 * ops + 32 * i + j]
                c.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                if j < 16:
                    clo = clo * 4 + limb
                else:
                    chi = chi * 4 + limb
            c.append(clo - lo)
            c.append(chi - hi)
    elif kind == GATE_U32_ADD_MANY:  # add_many_u32.rs:87-135
        na, ops, per = p0, p1, p0 + 3
        for i in range(ops):
            computed = w[0] * 0
            for j in range(na):
                computed = computed + w[per * i + j]
            computed = computed + w[per * i + na]
            res, carry = w[per * i + na + 1], w[per * i + na + 2]
            c.append(carry * (1 << 32) + res - computed)
            cres, ccar = w[0] * 0, w[0] * 0
            for j in reversed(range(18)):
                limb = w[per * ops + 18 * i + j]
                c.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                if j < 16:
                    cres = cres * 4 + limb
                else:
                    ccar = ccar * 4 + limb
            c.append(cres - res)
            c.append(ccar - carry)
    elif kind == GATE_U32_SUBTRACTION:  # subtraction_u32.rs:82-125
        ops = p0
        for i in range(ops):
            x, y, bin_, res, bout = (w[5 * i + k] for k in range(5))
            c.append(res - (x - y - bin_ + bout * (1 << 32)))
            comb = w[0] * 0
            for j in reversed(range(16)):
                limb = w[5 * ops + 16 * i + j]
                c.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                comb = comb * 4 + limb
            c.append(comb - res)
            c.append(bout * (1 - bout))
    elif kind == GATE_U32_RANGE_CHECK:  # range_check_u32.rs:51-75
        nl = p0
        for i in range(nl):
            comb = w[0] * 0
            for j in reversed(range(16)):
                comb = comb * 4 + w[nl + 16 * i + j]
            c.append(comb - w[i])
            for j in range(16):
                limb = w[nl + 16 * i + j]
                c.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
    elif kind == GATE_U32_INTERLEAVE:  # interleave_u32.rs:86-127; bits big-endian
        ops = p0
        for i in range(ops):
            x, xi = w[2 * i], w[2 * i + 1]
            bits = [w[2 * ops + 32 * i + j] for j in range(32)]
            cx, cxi = w[0] * 0, w[0] * 0
            for b in bits:  # reduce_with_powers(bits.rev(), base): Horner from the most significant bit
                cx = cx * 2 + b
                cxi = cxi * 4 + b
            c.append(cx - x)
            c.append(cxi - xi)
            for b in bits:
                c.append(b * (b - 1))
    elif kind in (GATE_UNINTERLEAVE_TO_U32, GATE_UNINTERLEAVE_TO_B32):  # uninterleave_to_u32.rs:93-136 / _b32.rs:97-141
        ops = p0
        for i in range(ops):
            xi, xe, xo = w[3 * i], w[3 * i + 1], w[3 * i + 2]
            bits = [w[3 * ops + 64 * i + j] for j in range(64)]
            cx = w[0] * 0
            for b in bits:
                cx = cx * 2 + b
            c.append(cx - xi)
            ce, co = w[0] * 0, w[0] * 0
            for j in range(32):
                coeff = (1 << (31 - j)) if kind == GATE_UNINTERLEAVE_TO_U32 else (1 << (2 * (31 - j)))
                ce = ce + bits[2 * j] * coeff
                co = co + bits[2 * j + 1] * coeff
            c.append(ce - xe)
            c.append(co - xo)
            for b in bits:
                c.append(b * (b - 1))
    elif kind == GATE_COMPARISON:  # comparison.rs:96-170; p0 = num_bits, p1 = num_chunks
        nb, nc = p0, p1
        cb = -(-nb // nc)
        first, second = w[0], w[1]
        fc = [w[4 + i] for i in range(nc)]
        sc = [w[4 + nc + i] for i in range(nc)]
        f_comb, s_comb = w[0] * 0, w[0] * 0
        for i in reversed(range(nc)):
            f_comb = f_comb * (1 << cb) + fc[i]
            s_comb = s_comb * (1 << cb) + sc[i]
        c.append(f_comb - first)
        c.append(s_comb - second)
        msd = w[0] * 0
        for i in range(nc):
            fp = fc[i] * 0 + 1
            sp = fp
            for x in range(1 << cb):
                fp = fp * (fc[i] - x)
                sp = sp * (sc[i] - x)
            c.append(fp)
            c.append(sp)
            diff = sc[i] - fc[i]
            dummy, eq = w[4 + 2 * nc + i], w[4 + 3 * nc + i]
            c.append(diff * dummy - (1 - eq))
            c.append(eq * diff)
            inter = w[4 + 4 * nc + i]
            c.append(inter - eq * msd)
            msd = inter + (1 - eq) * diff
        c.append(w[3] - msd)
        bits = [w[4 + 5 * nc + k] for k in range(cb + 1)]
        for b in bits:
            c.append(b * (1 - b))
        comb = w[0] * 0
        for b in reversed(bits):
            comb = comb * 2 + b
        c.append(w[3] + (1 << cb) - comb)
        c.append(w[2] - bits[cb])
    elif kind == GATE_ARITHMETIC_EXT:  # gates/arithmetic_extension.rs; 4 D wires per op
        for i in range(p0):
            m0, m1, ad, out = (_ext_at(w, 8 * i + 2 * k) for k in range(4))
            c.extend((out - (m0 * m1 * consts[0] + ad * consts[1])).parts())
    elif kind == GATE_MUL_EXT:  # gates/multiplication_extension.rs; 3 D wires per op
        for i in range(p0):
            m0, m1, out = (_ext_at(w, 6 * i + 2 * k) for k in range(3))
            c.extend((out - m0 * m1 * consts[0]).parts())
    elif kind in (GATE_REDUCING, GATE_REDUCING_EXT):  # gates/reducing.rs, reducing_extension.rs
        n = p0
        ext = kind == GATE_REDUCING_EXT
        alpha, acc = _ext_at(w, 2), _ext_at(w, 4)
        start_accs = 6 + (2 * n if ext else n)
        for i in range(n):
            coeff = _ext_at(w, 6 + 2 * i) if ext else Alg(w[6 + i], w[0] * 0)
            nxt = _ext_at(w, 0) if i == n - 1 else _ext_at(w, start_accs + 2 * i)
            c.extend((acc * alpha + coeff - nxt).parts())
            acc = nxt
    elif kind == GATE_RANDOM_ACCESS:  # gates/random_access.rs; p0 = bits, p1 = num_copies | num_extra_constants << 16
        bits, copies, extra = p0, p1 & 0xFFFF, p1 >> 16
        vec = 1 << bits
        routed = (2 + vec) * copies + extra
        for cp in range(copies):
            base = (2 + vec) * cp
            b = [w[routed + cp * bits + k] for k in range(bits)]
            for x in b:
                c.append(x * (x - 1))
            rec = w[0] * 0
            for x in reversed(b):
                rec = rec * 2 + x
            c.append(rec - w[base])
            items = [w[base + 2 + k] for k in range(vec)]
            for x in b:
                items = [items[2 * k] + x * (items[2 * k + 1] - items[2 * k]) for k in range(len(items) // 2)]
            c.append(items[0] - w[base + 1])
        for k in range(extra):
            c.append(consts[k] - w[(2 + vec) * copies + k])
    elif kind == GATE_POSEIDON_MDS:  # gates/poseidon_mds.rs; inputs 12 x D, outputs 12 x D
        ins = [_ext_at(w, 2 * i) for i in range(12)]
        for r in range(12):
            acc = ins[r] * 0
            for i in range(12):
                acc = acc + ins[(i + r) % 12] * PF.CIRC[i]
            acc = acc + ins[r] * PF.DIAG[r]
            c.extend((acc - _ext_at(w, 24 + 2 * r)).parts())
    elif kind == GATE_COSET_INTERPOLATION:  # gates/coset_interpolation.rs; p0 = subgroup_bits, p1 = degree
        c.extend(coset_interpolation_constraints(p0, p1, w))
    else:
        raise ValueError(kind)
    return c


def coset_interpolation_layout(bits, degree):
    n = 1 << bits
    n_int = (n - 2) // (degree - 1)
    start_point = 1 + 2 * n
    start_int = start_point + 4
    return dict(n=n, n_int=n_int, point=start_point, value=start_point + 2, inter_eval=start_int,
                inter_prod=start_int + 2 * n_int, shifted=start_int + 4 * n_int, end=start_int + 2 * (2 * n_int + 1))


def coset_interpolation_constraints(bits, degree, w):
    """barycentric interpolation over shift * H in chunks of `degree` (first) / `degree - 1` points; the weights of
    the subgroup H are x_i / n"""
    L = coset_interpolation_layout(bits, degree)
    n = L["n"]
    g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - bits), P)
    dom = [pow(g, i, P) for i in range(n)]
    ninv = pow(n, P - 2, P)
    wts = [x * ninv % P for x in dom]
    for i in range(n):  # the definition plonky2 uses: 1 / prod_{j != i} (x_i - x_j)
        if i < 3:
            d = 1
            for j in range(n):
                if j != i:
                    d = d * (dom[i] - dom[j]) % P
            assert wts[i] == pow(d, P - 2, P)
    shift = w[0]
    values = [_ext_at(w, 1 + 2 * i) for i in range(n)]
    point, shifted = _ext_at(w, L["point"]), _ext_at(w, L["shifted"])
    c = (point - shifted * shift).parts()
    one, zero = w[0] * 0 + 1, w[0] * 0

    def partial(lo, hi, ev, prod):
        for i in range(lo, hi):
            term = Alg(shifted.c0 - dom[i], shifted.c1)
            ev, prod = ev * term + values[i] * wts[i] * prod, prod * term
        return ev, prod

    ev, prod = partial(0, degree, Alg(zero, zero), Alg(one, zero))
    for i in range(L["n_int"]):
        ie, ip = _ext_at(w, L["inter_eval"] + 2 * i), _ext_at(w, L["inter_prod"] + 2 * i)
        c.extend((ie - ev).parts())
        c.extend((ip - prod).parts())
        lo = 1 + (degree - 1) * (i + 1)
        ev, prod = partial(lo, min(lo + degree - 1, n), ie, ip)
    c.extend((_ext_at(w, L["value"]) - ev).parts())
    return c


def gate_degree(kind):
    return {GATE_NOOP: 0, GATE_CONSTANT: 1, GATE_PUBLIC_INPUT: 1, GATE_ARITHMETIC: 3, GATE_POSEIDON: 7,
            GATE_BASE_SUM: 2, GATE_U32_ARITHMETIC: 4, GATE_U32_ADD_MANY: 4, GATE_U32_SUBTRACTION: 4,
            GATE_U32_RANGE_CHECK: 4, GATE_U32_INTERLEAVE: 2, GATE_UNINTERLEAVE_TO_U32: 2, GATE_UNINTERLEAVE_TO_B32: 2,
            GATE_COMPARISON: 4,  # ComparisonGate: 2^chunk_bits with chunk_bits = 2
            GATE_ARITHMETIC_EXT: 3, GATE_MUL_EXT: 3, GATE_REDUCING: 2, GATE_REDUCING_EXT: 2,
            GATE_RANDOM_ACCESS: 5,  # bits + 1 with bits = 4
            GATE_POSEIDON_MDS: 1, GATE_COSET_INTERPOLATION: 6}[kind]  # with_max_degree(4, 8) -> degree 6


def gate_num_constraints(kind, p0, p1):
    return {GATE_NOOP: 0, GATE_CONSTANT: p0, GATE_PUBLIC_INPUT: 4, GATE_ARITHMETIC: p0, GATE_POSEIDON: 123,
            GATE_BASE_SUM: 1 + p0, GATE_U32_ARITHMETIC: p0 * 36, GATE_U32_ADD_MANY: p1 * 21,
            GATE_U32_SUBTRACTION: p0 * 19, GATE_U32_RANGE_CHECK: p0 * 17, GATE_U32_INTERLEAVE: p0 * 34,
            GATE_UNINTERLEAVE_TO_U32: p0 * 67, GATE_UNINTERLEAVE_TO_B32: p0 * 67,
            GATE_COMPARISON: 6 + 5 * p1 + -(-p0 // max(p1, 1)), GATE_ARITHMETIC_EXT: 2 * p0, GATE_MUL_EXT: 2 * p0,
            GATE_REDUCING: 2 * p0, GATE_REDUCING_EXT: 2 * p0,
            GATE_RANDOM_ACCESS: (p0 + 2) * (p1 & 0xFFFF) + (p1 >> 16), GATE_POSEIDON_MDS: 24,
            GATE_COSET_INTERPOLATION: 4 + 4 * (((1 << p0) - 2) // max(p1 - 1, 1))}[kind]


# ------------------------------------------------------------------------------------------ witness generation
def poseidon_gate_wires(inputs, swap, rng):
    """wire values of one PoseidonGate row (plonky2 gates/poseidon.rs PoseidonGenerator, fast partial rounds)"""
    w = [rng.randrange(P) for _ in range(135)]
    w[0:12] = inputs
    w[24] = swap
    for i in range(4):
        w[25 + i] = swap * (inputs[i + 4] - inputs[i]) % P
    s = list(inputs)
    if swap:
        s[0:4], s[4:8] = inputs[4:8], inputs[0:4]
    rnd = 0
    for r in range(4):
        s = [(x + PF.RC[12 * rnd + i]) % P for i, x in enumerate(s)]
        if r != 0:
            w[29 + 12 * (r - 1):29 + 12 * r] = s
        s = PF.matvec(PF.M, [pow(x, 7, P) for x in s])
        rnd += 1
    s = [(x + k) % P for x, k in zip(s, PF.first_const)]
    s = [s[0]] + PF.matvec(PF.INIT, s[1:])
    for r in range(22):
        w[65 + r] = s[0]
        s[0] = pow(s[0], 7, P)
        if r < 21:
            s[0] = (s[0] + PF.post[r]) % P
        d = (s[0] * PF.M[0][0] + sum(a * b for a, b in zip(PF.W_HATS[r], s[1:]))) % P
        s = [d] + [(s[i] + s[0] * PF.VS[r][i - 1]) % P for i in range(1, 12)]
    rnd += 22
    for r in range(4):
        s = [(x + PF.RC[12 * rnd + i]) % P for i, x in enumerate(s)]
        w[87 + 12 * r:87 + 12 * (r + 1)] = s
        s = PF.matvec(PF.M, [pow(x, 7, P) for x in s])
        rnd += 1
    w[12:24] = s
    return w


def _limbs2(v, count):
    return [(v >> (2 * j)) & 3 for j in range(count)]


class SyntheticCircuit:
    """Random satisfiable circuit over a list of gates [(kind, p0, p1)], split into selector groups."""

    def __init__(self, degree_bits, gates, groups, seed, num_wires=135, num_routed=80, num_gate_consts=2,
                 num_challenges=2, quotient_degree_factor=8, link_prob=0.3, pi_hash=None):
        rng = random.Random(seed)
        self.degree_bits, self.n = degree_bits, 1 << degree_bits
        n = self.n
        self.gates, self.groups = gates, groups
        self.num_wires, self.num_routed = num_wires, num_routed
        self.num_selectors = len(groups)
        self.num_constants = self.num_selectors + num_gate_consts
        self.num_challenges, self.qdf = num_challenges, quotient_degree_factor
        self.num_pp = -(-num_routed // quotient_degree_factor) - 1
        self.k_is = [pow(7, j, P) for j in range(num_routed)]
        self.pi_hash = [rng.randrange(P) for _ in range(4)] if pi_hash is None else [int(x) for x in pi_hash]
        sel_of = {}
        for gi, (a, b) in enumerate(groups):
            for g in range(a, b):
                sel_of[g] = gi
        self.sel_of = sel_of
        # plonky2 groups gates so that selector filter + gate stay within the quotient degree bound
        for a, b in groups:
            filt = (b - a - 1) + (1 if len(groups) > 1 else 0)
            assert all(filt + gate_degree(gates[g][0]) <= quotient_degree_factor for g in range(a, b)), "degree bound"
        # every gate at least once (public input first, as plonky2 places it), then random
        order = list(range(len(gates)))
        row_gate = order + [rng.randrange(len(gates)) for _ in range(n - len(order))]
        assert len(row_gate) == n
        self.row_gate = row_gate
        consts = [[0] * n for _ in range(self.num_constants)]
        wires = [[0] * n for _ in range(num_wires)]
        parent = {}

        def find(x):
            while parent.get(x, x) != x:
                parent[x] = parent.get(parent[x], parent[x])
                x = parent[x]
            return x

        def pick(row, col):
            """value for a free felt-typed routed input: maybe copied from an earlier routed slot"""
            if row > 0 and col < num_routed and rng.random() < link_prob:
                r2, c2 = rng.randrange(row), rng.randrange(num_routed)
                parent[find((row, col))] = find((r2, c2))
                return wires[c2][r2]
            return rng.randrange(P)

        for i, g in enumerate(row_gate):
            kind, p0, p1 = gates[g]
            for s in range(self.num_selectors):
                consts[s][i] = g if s == sel_of[g] else UNUSED_SELECTOR
            gc = [rng.randrange(P) for _ in range(num_gate_consts)]
            row = [None] * num_wires
            if kind == GATE_NOOP:
                row = [pick(i, c) for c in range(num_wires)]
            elif kind == GATE_CONSTANT:
                row = [gc[c] if c < p0 else pick(i, c) for c in range(num_wires)]
            elif kind == GATE_PUBLIC_INPUT:
                row = [self.pi_hash[c] if c < 4 else pick(i, c) for c in range(num_wires)]
            elif kind == GATE_ARITHMETIC:
                row = [0 if (c < 4 * p0 and c % 4 == 3) else pick(i, c) for c in range(num_wires)]
                for o in range(p0):
                    row[4 * o + 3] = (row[4 * o] * row[4 * o + 1] * gc[0] + row[4 * o + 2] * gc[1]) % P
            elif kind == GATE_POSEIDON:
                inputs = [pick(i, c) for c in range(12)]
                row = poseidon_gate_wires(inputs, rng.randrange(2), rng)
            elif kind == GATE_BASE_SUM:
                row = [rng.randrange(P) for _ in range(num_wires)]
                bits = [rng.randrange(2) for _ in range(p0)]
                row[1:1 + p0] = bits
                row[0] = sum(b << k for k, b in enumerate(bits))
            elif kind == GATE_U32_ARITHMETIC:
                row = [rng.randrange(P) for _ in range(num_wires)]
                for o in range(p0):
                    m0, m1, ad = (rng.randrange(2**32) for _ in range(3))
                    if o == 0:
                        m0 = m1 = ad = 2**32 - 1  # high limb = u32::MAX - 1, edge of the canonicity check
                    out = m0 * m1 + ad
                    lo, hi = out & 0xFFFFFFFF, out >> 32
                    inv = pow((0xFFFFFFFF - hi) % P, P - 2, P)
                    row[6 * o:6 * o + 6] = [m0, m1, ad, lo, hi, inv]
                    row[6 * p0 + 32 * o:6 * p0 + 32 * (o + 1)] = _limbs2(lo, 16) + _limbs2(hi, 16)
            elif kind == GATE_U32_ADD_MANY:
                row = [rng.randrange(P) for _ in range(num_wires)]
                na, ops, per = p0, p1, p0 + 3
                for o in range(ops):
                    add = [rng.randrange(2**32) for _ in range(na)]
                    carry = rng.randrange(2**32)
                    tot = sum(add) + carry
                    res, oc = tot & 0xFFFFFFFF, tot >> 32
                    row[per * o:per * (o + 1)] = add + [carry, res, oc]
                    row[per * ops + 18 * o:per * ops + 18 * (o + 1)] = _limbs2(res, 16) + _limbs2(oc, 2)
            elif kind == GATE_U32_SUBTRACTION:
                row = [rng.randrange(P) for _ in range(num_wires)]
                for o in range(p0):
                    x, y, bi = rng.randrange(2**32), rng.randrange(2**32), rng.randrange(2)
                    d = x - y - bi
                    bo = 1 if d < 0 else 0
                    res = d + (bo << 32)
                    row[5 * o:5 * o + 5] = [x, y, bi, res, bo]
                    row[5 * p0 + 16 * o:5 * p0 + 16 * (o + 1)] = _limbs2(res, 16)
            elif kind == GATE_U32_RANGE_CHECK:
                row = [rng.randrange(P) for _ in range(num_wires)]
                for o in range(p0):
                    v = rng.randrange(2**32)
                    row[o] = v
                    row[p0 + 16 * o:p0 + 16 * (o + 1)] = _limbs2(v, 16)
            elif kind == GATE_U32_INTERLEAVE:
                row = [rng.randrange(P) for _ in range(num_wires)]
                for o in range(p0):
                    x = rng.randrange(2**32)
                    bits = [(x >> (31 - j)) & 1 for j in range(32)]
                    row[2 * o] = x
                    row[2 * o + 1] = sum(b << (2 * (31 - j)) for j, b in enumerate(bits))
                    row[2 * p0 + 32 * o:2 * p0 + 32 * (o + 1)] = bits
            elif kind in (GATE_UNINTERLEAVE_TO_U32, GATE_UNINTERLEAVE_TO_B32):
                row = [rng.randrange(P) for _ in range(num_wires)]
                for o in range(p0):
                    x = rng.randrange(2**63)
                    bits = [(x >> (63 - j)) & 1 for j in range(64)]
                    step = 1 if kind == GATE_UNINTERLEAVE_TO_U32 else 2
                    ev = sum(bits[2 * j] << (step * (31 - j)) for j in range(32))
                    od = sum(bits[2 * j + 1] << (step * (31 - j)) for j in range(32))
                    row[3 * o:3 * o + 3] = [x, ev, od]
                    row[3 * p0 + 64 * o:3 * p0 + 64 * (o + 1)] = bits
            elif kind == GATE_COMPARISON:
                row = [rng.randrange(P) for _ in range(num_wires)]
                nb, nc = p0, p1
                cb = -(-nb // nc)
                a, b = rng.randrange(2**nb), rng.randrange(2**nb)
                if i % 3 == 0:
                    b = a  # equal inputs
                elif i % 3 == 1:
                    b = (a & ~0xFF) | (b & 0xFF)  # equal high chunks
                fc = [(a >> (cb * k)) & ((1 << cb) - 1) for k in range(nc)]
                sc = [(b >> (cb * k)) & ((1 << cb) - 1) for k in range(nc)]
                msd = 0
                row[0], row[1] = a, b
                for k in range(nc):
                    diff = (sc[k] - fc[k]) % P
                    eq = 1 if diff == 0 else 0
                    inter = eq * msd % P
                    row[4 + k], row[4 + nc + k] = fc[k], sc[k]
                    row[4 + 2 * nc + k] = 1 if eq else pow(diff, P - 2, P)
                    row[4 + 3 * nc + k] = eq
                    row[4 + 4 * nc + k] = inter
                    msd = (inter + (1 - eq) * diff) % P
                row[3] = msd
                comb = ((1 << cb) + msd) % P
                bits = [(comb >> k) & 1 for k in range(cb + 1)]
                row[4 + 5 * nc:4 + 5 * nc + cb + 1] = bits
                row[2] = bits[cb]
            elif kind in (GATE_ARITHMETIC_EXT, GATE_MUL_EXT):
                per = 8 if kind == GATE_ARITHMETIC_EXT else 6
                row = [pick(i, c) if c < per * p0 and c % per < per - 2 else rng.randrange(P) for c in range(num_wires)]
                for o in range(p0):
                    m0, m1 = Ext(row[per * o], row[per * o + 1]), Ext(row[per * o + 2], row[per * o + 3])
                    out = m0 * m1 * gc[0]
                    if kind == GATE_ARITHMETIC_EXT:
                        out = out + Ext(row[per * o + 4], row[per * o + 5]) * gc[1]
                    row[per * o + per - 2], row[per * o + per - 1] = out.a, out.b
            elif kind in (GATE_REDUCING, GATE_REDUCING_EXT):
                row = [rng.randrange(P) for _ in range(num_wires)]
                ext = kind == GATE_REDUCING_EXT
                alpha, acc = Ext(row[2], row[3]), Ext(row[4], row[5])
                start_accs = 6 + (2 * p0 if ext else p0)
                for k in range(p0):
                    coeff = Ext(row[6 + 2 * k], row[7 + 2 * k]) if ext else Ext(row[6 + k])
                    acc = acc * alpha + coeff
                    pos = 0 if k == p0 - 1 else start_accs + 2 * k
                    row[pos], row[pos + 1] = acc.a, acc.b
            elif kind == GATE_RANDOM_ACCESS:
                row = [rng.randrange(P) for _ in range(num_wires)]
                bits, copies, extra = p0, p1 & 0xFFFF, p1 >> 16
                vec = 1 << bits
                routed = (2 + vec) * copies + extra
                for cp in range(copies):
                    base = (2 + vec) * cp
                    idx = rng.randrange(vec)
                    row[base] = idx
                    row[base + 1] = row[base + 2 + idx]
                    row[routed + cp * bits:routed + (cp + 1) * bits] = [(idx >> k) & 1 for k in range(bits)]
                for k in range(extra):
                    row[(2 + vec) * copies + k] = gc[k]
            elif kind == GATE_POSEIDON_MDS:
                row = [pick(i, c) if c < 24 else rng.randrange(P) for c in range(num_wires)]
                for part in range(2):
                    v = PF.matvec(PF.M, [row[2 * k + part] for k in range(12)])
                    for r in range(12):
                        row[24 + 2 * r + part] = v[r]
            elif kind == GATE_COSET_INTERPOLATION:
                row = [rng.randrange(P) for _ in range(num_wires)]
                L = coset_interpolation_layout(p0, p1)
                npts = L["n"]
                g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - p0), P)
                dom = [pow(g, k, P) for k in range(npts)]
                ninv = pow(npts, P - 2, P)
                shift = row[0] or 1
                row[0] = shift
                point = Ext(row[L["point"]], row[L["point"] + 1])
                shifted = point * pow(shift, P - 2, P)
                row[L["shifted"]], row[L["shifted"] + 1] = shifted.a, shifted.b
                vals = [Ext(row[1 + 2 * k], row[2 + 2 * k]) for k in range(npts)]
                ev, prod = Ext(0), Ext(1)

                def part(lo, hi, ev, prod):
                    for k in range(lo, hi):
                        term = shifted - dom[k]
                        ev, prod = ev * term + vals[k] * (dom[k] * ninv % P) * prod, prod * term
                    return ev, prod

                ev, prod = part(0, p1, ev, prod)
                for k in range(L["n_int"]):
                    row[L["inter_eval"] + 2 * k], row[L["inter_eval"] + 2 * k + 1] = ev.a, ev.b
                    row[L["inter_prod"] + 2 * k], row[L["inter_prod"] + 2 * k + 1] = prod.a, prod.b
                    lo = 1 + (p1 - 1) * (k + 1)
                    ev, prod = part(lo, min(lo + p1 - 1, npts), ev, prod)
                row[L["value"]], row[L["value"] + 1] = ev.a, ev.b
                # semantic check: the gate's output is the Lagrange interpolant through (shift x_k, a_k) at `point`
                lag = Ext(0)
                for k in range(npts):
                    num, den = Ext(1), 1
                    for j in range(npts):
                        if j != k:
                            num = num * (point - shift * dom[j])
                            den = den * (shift * (dom[k] - dom[j])) % P
                    lag = lag + vals[k] * num * pow(den, P - 2, P)
                assert lag == ev, "CosetInterpolationGate restatement does not interpolate"
            for c in range(num_gate_consts):
                consts[self.num_selectors + c][i] = gc[c]
            for c in range(num_wires):
                wires[c][i] = row[c]
            # the gate's own constraints hold on this row
            cons = eval_gate(kind, p0, p1, [Fp(x) for x in row], [Fp(x) for x in gc], [Fp(x) for x in self.pi_hash])
            assert all(x == 0 for x in cons), (kind, [k for k, x in enumerate(cons) if not x == 0][:5])
        self.wires, self.consts = wires, consts
        # sigma: every class of linked slots becomes one cycle
        classes = {}
        for r in range(n):
            for c in range(num_routed):
                classes.setdefault(find((r, c)), []).append((r, c))
        omega = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - degree_bits), P)
        self.subgroup = [pow(omega, i, P) for i in range(n)]
        sigmas = [[0] * n for _ in range(num_routed)]
        for members in classes.values():
            assert len({wires[c][r] for r, c in members}) == 1
            for idx, (r, c) in enumerate(members):
                r2, c2 = members[(idx + 1) % len(members)]
                sigmas[c][r] = self.k_is[c2] * self.subgroup[r2] % P
        self.sigmas = sigmas
        self.num_gate_constraints = max(gate_num_constraints(*g) for g in gates)

    def desc(self):
        gates = []
        for g, (kind, p0, p1) in enumerate(self.gates):
            s = self.sel_of[g]
            gates.append(dict(kind=kind, p0=p0, p1=p1, selector_index=s, group_start=self.groups[s][0],
                              group_end=self.groups[s][1], row=g))
        return dict(degree_bits=self.degree_bits, num_wires=self.num_wires, num_routed_wires=self.num_routed,
                    num_constants=self.num_constants, num_selectors=self.num_selectors,
                    num_challenges=self.num_challenges, quotient_degree_factor=self.qdf,
                    num_partial_products=self.num_pp, num_gate_constraints=self.num_gate_constraints,
                    gates=gates, k_is=self.k_is)

    def constants_sigmas_values(self):
        return [np.array(c, dtype=np.uint64) for c in self.consts + self.sigmas]

    def wire_values(self):
        return [np.array(c, dtype=np.uint64) for c in self.wires]


# ------------------------------------------------------------------------------------------ verifier identity
def horner_ext(coeffs, x):
    acc = Ext(0)
    for c in reversed(coeffs):
        acc = acc * x + int(c)
    return acc


def eval_vanishing_poly_ext(circ, zeta, consts_z, sigmas_z, wires_z, zs_z, zs_next_z, pps_z, betas, gammas, alphas):
    """plonky2 plonk/vanishing_poly.rs::eval_vanishing_poly at zeta (no lookups); returns one Ext per challenge.
    pps_z[i] = the num_pp partial-product openings of challenge i."""
    n = circ.n
    zeta_n = zeta ** n
    z_h = zeta_n - 1
    l0 = z_h * (Ext(n) * (zeta - 1)).inv()
    terms_z1, terms_pp = [], []
    deg = circ.qdf
    for i in range(circ.num_challenges):
        terms_z1.append(l0 * (zs_z[i] - 1))
        num = [wires_z[j] + zeta * (betas[i] * circ.k_is[j] % P) + gammas[i] for j in range(circ.num_routed)]
        den = [wires_z[j] + sigmas_z[j] * betas[i] + gammas[i] for j in range(circ.num_routed)]
        accs = [zs_z[i]] + list(pps_z[i]) + [zs_next_z[i]]
        for k in range(circ.num_pp + 1):
            np_, dp = Ext(1), Ext(1)
            for j in range(k * deg, min((k + 1) * deg, circ.num_routed)):
                np_, dp = np_ * num[j], dp * den[j]
            terms_pp.append(accs[k] * np_ - accs[k + 1] * dp)
    gate_terms = [Ext(0)] * circ.num_gate_constraints
    pi = [Ext(x) for x in circ.pi_hash]
    for g, (kind, p0, p1) in enumerate(circ.gates):
        s = consts_z[circ.sel_of[g]]
        a, b = circ.groups[circ.sel_of[g]]
        filt = Ext(1)
        for i in range(a, b):
            if i != g:
                filt = filt * (Ext(i) - s)
        if circ.num_selectors > 1:
            filt = filt * (Ext(UNUSED_SELECTOR) - s)
        cons = eval_gate(kind, p0, p1, wires_z, consts_z[circ.num_selectors:], pi)
        for k, cv in enumerate(cons):
            gate_terms[k] = gate_terms[k] + cv * filt
    terms = terms_z1 + terms_pp + gate_terms
    out = []
    for i in range(circ.num_challenges):
        acc = Ext(0)
        for tv in reversed(terms):
            acc = acc * alphas[i] + tv
        out.append(acc)
    return out, z_h, zeta_n


# ------------------------------------------------------------------------------------------ gate sets of the test / bench circuits
ALL_GATES = [(GATE_PUBLIC_INPUT, 0, 0), (GATE_NOOP, 0, 0), (GATE_CONSTANT, 2, 0), (GATE_ARITHMETIC, 20, 0),
             (GATE_POSEIDON, 0, 0), (GATE_BASE_SUM, 63, 0), (GATE_U32_ARITHMETIC, 3, 0),
             (GATE_U32_ADD_MANY, 3, 5), (GATE_U32_SUBTRACTION, 6, 0), (GATE_U32_RANGE_CHECK, 7, 0)]
# the in-tree bit-manipulation / comparison gates, with the parameters the reference registers
# (city_common_circuit/src/builder/pad_circuit.rs:31-55: ComparisonGate::new(32, 16))
MORE_GATES = [(GATE_NOOP, 0, 0), (GATE_U32_INTERLEAVE, 3, 0), (GATE_UNINTERLEAVE_TO_U32, 2, 0),
              (GATE_UNINTERLEAVE_TO_B32, 2, 0), (GATE_COMPARISON, 32, 16), (GATE_POSEIDON, 0, 0)]
MORE_GROUPS = [(0, 3), (3, 5), (5, 6)]
# upstream extension-field gates of the recursion gate set, with the parameters standard_recursion_config gives
# them (ReducingGate(43) / ReducingExtensionGate(32) / RandomAccessGate(bits 4): builder/pad_circuit.rs:31-55)
EXT_GATES = [(GATE_NOOP, 0, 0), (GATE_ARITHMETIC_EXT, 10, 0), (GATE_MUL_EXT, 13, 0), (GATE_REDUCING, 43, 0),
             (GATE_REDUCING_EXT, 32, 0), (GATE_RANDOM_ACCESS, 4, 4 | (2 << 16)), (GATE_POSEIDON_MDS, 0, 0)]
EXT_GROUPS = [(0, 3), (3, 6), (6, 7)]
# the 13 gate types of plonky2's standard recursion circuits = the gate set of the proofs stored in
# qbench_data/example.bin (135 wires, num_gate_constraints 123: zk_signature2/mod.rs:54-57)
RECURSION_GATES = [(GATE_NOOP, 0, 0), (GATE_CONSTANT, 2, 0), (GATE_PUBLIC_INPUT, 0, 0), (GATE_BASE_SUM, 63, 0),
                   (GATE_REDUCING_EXT, 32, 0), (GATE_REDUCING, 43, 0), (GATE_ARITHMETIC_EXT, 10, 0),
                   (GATE_ARITHMETIC, 20, 0), (GATE_MUL_EXT, 13, 0), (GATE_POSEIDON_MDS, 0, 0),
                   (GATE_RANDOM_ACCESS, 4, 4 | (2 << 16)), (GATE_COSET_INTERPOLATION, 4, 6), (GATE_POSEIDON, 0, 0)]
RECURSION_GROUPS = [(0, 6), (6, 10), (10, 12), (12, 13)]
# The gate set a City Rollup op circuit carries: add_city_common_gates (city_common_circuit/src/builder/pad_circuit.rs:31-55:
# Constant, Comparison(32, 16), RandomAccess(4), Poseidon, PoseidonMds, Reducing(43), ReducingExtension(32), Arithmetic,
# ArithmeticExtension, MulExtension, BaseSum<2>, + the coset gate) next to Noop / PublicInput, plus the seven other in-tree
# u32 gates its gadgets add (city_common_circuit/src/u32/gates/*.rs) — all 21 gate kinds.  Selector groups as plonky2 forms
# them: gates sorted by degree, packed greedily while group size + max gate degree <= 8.
CITY_GATES = [(GATE_NOOP, 0, 0), (GATE_CONSTANT, 2, 0), (GATE_PUBLIC_INPUT, 0, 0), (GATE_POSEIDON_MDS, 0, 0),
              (GATE_BASE_SUM, 63, 0), (GATE_REDUCING_EXT, 32, 0),
              (GATE_REDUCING, 43, 0), (GATE_U32_INTERLEAVE, 3, 0), (GATE_UNINTERLEAVE_TO_U32, 2, 0),
              (GATE_UNINTERLEAVE_TO_B32, 2, 0), (GATE_ARITHMETIC_EXT, 10, 0),
              (GATE_ARITHMETIC, 20, 0), (GATE_MUL_EXT, 13, 0), (GATE_COMPARISON, 32, 16), (GATE_U32_ARITHMETIC, 3, 0),
              (GATE_U32_ADD_MANY, 3, 5), (GATE_U32_SUBTRACTION, 6, 0), (GATE_U32_RANGE_CHECK, 7, 0),
              (GATE_RANDOM_ACCESS, 4, 4 | (2 << 16)), (GATE_COSET_INTERPOLATION, 4, 6),
              (GATE_POSEIDON, 0, 0)]
CITY_GROUPS = [(0, 6), (6, 11), (11, 15), (15, 18), (18, 20), (20, 21)]
# CosetInterpolationGate::with_max_degree(4, max_quotient_degree_factor = 8): degree 6, two intermediates
COSET_GATES = [(GATE_NOOP, 0, 0), (GATE_COSET_INTERPOLATION, 4, 6), (GATE_CONSTANT, 2, 0), (GATE_ARITHMETIC, 20, 0)]
COSET_GROUPS = [(0, 2), (2, 4)]
