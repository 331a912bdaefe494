"""BLS12-381 pairing for the oracle's verify_proof (pairing/src/bls12_381/{fq6,fq12,mod}.rs restated generically).

ORACLE -- test infrastructure only (see oracle/__init__.py).  Slow and obvious on purpose: Fq12 = Fq6[w]/(w^2 - v),
Fq6 = Fq2[v]/(v^3 - (u + 1)) (fq6.rs, fq12.rs), an affine Miller loop over the untwisted G2 point and the reference's
final exponentiation chain (mod.rs:103-160) with Frobenius maps done as plain powers.  Pinned by the reference's RELIC
known-answer test (pairing/src/bls12_381/tests/mod.rs:5-53, tests/golden/kat.json["pairing_g1_g2"]).
"""
from __future__ import annotations

from .fields import FQ_MODULUS as Q, Fq, Fq2
from .curve import G1, G2

XI = (1, 1)  # u + 1, the cubic / sextic non-residue (fq2.rs:51-58)
BLS_X = 0xd201000000010000  # mod.rs:24-25, negative

F2 = Fq2
ZERO2, ONE2 = (0, 0), (1, 0)


# ---- Fq6: (c0, c1, c2) = c0 + c1 v + c2 v^2, v^3 = xi   (fq6.rs)
def f6_add(a, b): return tuple(F2.add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(F2.sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(F2.neg(x) for x in a)
def f6_mul(a, b):
    a0, a1, a2 = a; b0, b1, b2 = b
    m = F2.mul
    c0 = F2.add(m(a0, b0), m(XI, F2.add(m(a1, b2), m(a2, b1))))
    c1 = F2.add(F2.add(m(a0, b1), m(a1, b0)), m(XI, m(a2, b2)))
    c2 = F2.add(F2.add(m(a0, b2), m(a1, b1)), m(a2, b0))
    return (c0, c1, c2)
def f6_mul_by_v(a): return (F2.mul(XI, a[2]), a[0], a[1])
def f6_inv(a):
    a0, a1, a2 = a
    m = F2.mul
    t0 = F2.sub(F2.sqr(a0), m(XI, m(a1, a2)))
    t1 = F2.sub(m(XI, F2.sqr(a2)), m(a0, a1))
    t2 = F2.sub(F2.sqr(a1), m(a0, a2))
    d = F2.add(m(a0, t0), m(XI, F2.add(m(a2, t1), m(a1, t2))))
    di = F2.inv(d)
    return (m(t0, di), m(t1, di), m(t2, di))
ZERO6 = (ZERO2, ZERO2, ZERO2)
ONE6 = (ONE2, ZERO2, ZERO2)


# ---- Fq12: (c0, c1) = c0 + c1 w, w^2 = v   (fq12.rs)
def f12_mul(a, b):
    a0, a1 = a; b0, b1 = b
    return (f6_add(f6_mul(a0, b0), f6_mul_by_v(f6_mul(a1, b1))), f6_add(f6_mul(a0, b1), f6_mul(a1, b0)))
def f12_sqr(a): return f12_mul(a, a)
def f12_add(a, b): return (f6_add(a[0], b[0]), f6_add(a[1], b[1]))
def f12_sub(a, b): return (f6_sub(a[0], b[0]), f6_sub(a[1], b[1]))
def f12_conj(a): return (a[0], f6_neg(a[1]))
def f12_inv(a):
    a0, a1 = a
    d = f6_sub(f6_mul(a0, a0), f6_mul_by_v(f6_mul(a1, a1)))
    di = f6_inv(d)
    return (f6_mul(a0, di), f6_neg(f6_mul(a1, di)))
ONE12 = (ONE6, ZERO6)
def f12_pow(a, e):
    r = ONE12
    for i in reversed(range(e.bit_length())):
        r = f12_sqr(r)
        if (e >> i) & 1:
            r = f12_mul(r, a)
    return r
def f12_frobenius(a, k): return f12_pow(a, Q ** k)
def f12_from_fq(x): return (((x % Q, 0), ZERO2, ZERO2), ZERO6)
def f12_from_fq2(x): return ((x, ZERO2, ZERO2), ZERO6)
W = (ZERO6, ONE6)                      # w
W2 = f12_mul(W, W)                     # w^2 = v
W3 = f12_mul(W2, W)
W2_INV, W3_INV = f12_inv(W2), f12_inv(W3)


def untwist(q_aff):
    """psi: E'(Fq2) -> E(Fq12), (x, y) -> (x / w^2, y / w^3)  (M-type twist y^2 = x^3 + 4(u+1), w^6 = u+1)"""
    return (f12_mul(f12_from_fq2(q_aff[0]), W2_INV), f12_mul(f12_from_fq2(q_aff[1]), W3_INV))


def miller_loop(p_aff, q_aff):
    """f_{|x|, Q}(P) with affine line functions over Fq12, conjugated because x < 0 (mod.rs:47-101)."""
    if p_aff[2] or q_aff[2]:
        return ONE12
    xp, yp = f12_from_fq(p_aff[0]), f12_from_fq(p_aff[1])
    xq, yq = untwist(q_aff)
    xt, yt = xq, yq
    f = ONE12
    three = f12_from_fq(3)
    two = f12_from_fq(2)

    def line(x1, y1, lam):
        return f12_sub(f12_sub(yp, y1), f12_mul(lam, f12_sub(xp, x1)))

    for i in reversed(range(BLS_X.bit_length() - 1)):
        lam = f12_mul(f12_mul(three, f12_sqr(xt)), f12_inv(f12_mul(two, yt)))
        f = f12_mul(f12_sqr(f), line(xt, yt, lam))
        x3 = f12_sub(f12_sub(f12_sqr(lam), xt), xt)
        yt = f12_sub(f12_mul(lam, f12_sub(xt, x3)), yt)
        xt = x3
        if (BLS_X >> i) & 1:
            lam = f12_mul(f12_sub(yq, yt), f12_inv(f12_sub(xq, xt)))
            f = f12_mul(f, line(xt, yt, lam))
            x3 = f12_sub(f12_sub(f12_sqr(lam), xt), xq)
            yt = f12_sub(f12_mul(lam, f12_sub(xt, x3)), yt)
            xt = x3
    return f12_conj(f)


def final_exponentiation(r):
    """mod.rs:103-160, literally (the hard part is the reference's x-chain)."""
    f1 = f12_conj(r)
    f2 = f12_inv(r)
    r = f12_mul(f1, f2)
    f2 = r
    r = f12_mul(f12_frobenius(r, 2), f2)

    def exp_by_x(f, x):
        return f12_conj(f12_pow(f, x))  # BLS_X_IS_NEGATIVE

    x = BLS_X
    y0 = f12_sqr(r)
    y1 = exp_by_x(y0, x)
    x >>= 1
    y2 = exp_by_x(y1, x)
    x <<= 1
    y3 = f12_conj(r)
    y1 = f12_mul(y1, y3)
    y1 = f12_conj(y1)
    y1 = f12_mul(y1, y2)
    y2 = exp_by_x(y1, x)
    y3 = exp_by_x(y2, x)
    y1 = f12_conj(y1)
    y3 = f12_mul(y3, y1)
    y1 = f12_conj(y1)
    y1 = f12_frobenius(y1, 3)
    y2 = f12_frobenius(y2, 2)
    y1 = f12_mul(y1, y2)
    y2 = exp_by_x(y3, x)
    y2 = f12_mul(y2, y0)
    y2 = f12_mul(y2, r)
    y1 = f12_mul(y1, y2)
    y2 = f12_frobenius(y3, 1)
    y1 = f12_mul(y1, y2)
    return y1


def pairing(p_aff, q_aff):
    """Engine::pairing (pairing/src/lib.rs:86-96)"""
    return final_exponentiation(miller_loop(p_aff, q_aff))


def pairing_product_is_one(pairs):
    f = ONE12
    for p, q in pairs:
        f = f12_mul(f, miller_loop(p, q))
    return final_exponentiation(f) == ONE12


class Bls12:
    """The BLS12-381 engine for oracle.groth16 (Fr, G1, G2, pairing check)."""
    from .fields import Fr as _Fr
    Fr = _Fr
    G1 = G1
    G2 = G2
    pairing_product_is_one = staticmethod(pairing_product_is_one)


def f12_flat(a):
    """c0.c0.c0, c0.c0.c1, c0.c1.c0, ... as 12 ints (the order of the reference's KAT literal)"""
    return [c for six in a for two in six for c in two]
