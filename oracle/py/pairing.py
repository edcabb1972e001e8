"""Textbook ate pairing on BLS12-381 for the test oracle (TEST INFRASTRUCTURE
ONLY; see oracle/py/bls.py).

Deliberately written in a different style from the CUDA code it checks:
Fp12 is the flat quotient Fp[w]/(w^12 - 2 w^6 + 2) (w^6 = 1 + u), G2 points are
untwisted into E(Fp12) and the Miller loop uses plain affine chord-and-tangent
lines there; the final exponentiation is one big pow by (p^12 - 1)/r.

Restates lambdaworks-math BLS12381AtePairing::compute_batch as used by
KZG::verify (/root/reference/src/lib.rs:444,496,691);
only "product == 1" is observable.
"""
from __future__ import annotations

from .bls import BLS_X, P, R

DEG = 12


def _reduce(c):
    # w^12 = 2 w^6 - 2
    c = list(c)
    for i in range(len(c) - 1, DEG - 1, -1):
        v = c[i]
        if v:
            c[i - 6] = (c[i - 6] + 2 * v) % P
            c[i - 12] = (c[i - 12] - 2 * v) % P
    return tuple(c[:DEG])


def f12(*coeffs):
    c = list(coeffs) + [0] * (DEG - len(coeffs))
    return tuple(x % P for x in c)


ONE = f12(1)
ZERO = f12(0)
W_RAW = f12(0, 1)


def add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def mul(a, b):
    out = [0] * (2 * DEG - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] += x * y
    return _reduce([v % P for v in out])


def scalar(a, k):
    return tuple(x * k % P for x in a)


def fpow(a, e):
    r = ONE
    for bit in bin(e)[2:]:
        r = mul(r, r)
        if bit == "1":
            r = mul(r, a)
    return r


def inv(a):
    """Inverse by solving the 12x12 linear system (mul-by-a) x = 1 over Fp."""
    cols = []
    basis = a
    for _ in range(DEG):
        cols.append(basis)
        basis = mul(basis, W_RAW)
    # matrix M[row][col] = cols[col][row]; augmented with e_0
    m = [[cols[c][r] for c in range(DEG)] + [1 if r == 0 else 0] for r in range(DEG)]
    for col in range(DEG):
        piv = next(r for r in range(col, DEG) if m[r][col])
        m[col], m[piv] = m[piv], m[col]
        iv = pow(m[col][col], -1, P)
        m[col] = [v * iv % P for v in m[col]]
        for r in range(DEG):
            if r != col and m[r][col]:
                f = m[r][col]
                m[r] = [(x - f * y) % P for x, y in zip(m[r], m[col])]
    return tuple(m[r][DEG] for r in range(DEG))


def fp2_to_f12(a):
    """a0 + a1 u with u = w^6 - 1."""
    c = [0] * DEG
    c[0] = (a[0] - a[1]) % P
    c[6] = a[1] % P
    return tuple(c)


W = f12(0, 1)
W2 = mul(W, W)
W3 = mul(W2, W)
W2_INV = inv(W2)
W3_INV = inv(W3)


def untwist(q):
    """psi: E'(Fp2) -> E(Fp12), (x, y) -> (x / w^2, y / w^3)."""
    x, y = q
    return (mul(fp2_to_f12(x), W2_INV), mul(fp2_to_f12(y), W3_INV))


def _line(t, s, p):
    """Line through t and s (points of E(Fp12)) evaluated at p; returns (value, t+s)."""
    (x1, y1), (x2, y2) = t, s
    xp, yp = p
    if x1 != x2:
        lam = mul(sub(y2, y1), inv(sub(x2, x1)))
    elif y1 == y2:
        lam = mul(scalar(mul(x1, x1), 3), inv(scalar(y1, 2)))
    else:
        return sub(xp, x1), None
    x3 = sub(sub(mul(lam, lam), x1), x2)
    y3 = sub(mul(lam, sub(x1, x3)), y1)
    val = sub(sub(yp, y1), mul(lam, sub(xp, x1)))
    return val, (x3, y3)


def miller(p, q):
    """f_{|x|,Q}(P), conjugated for the negative BLS parameter."""
    if p is None or q is None:
        return ONE
    pp = (f12(p[0]), f12(p[1]))
    qq = untwist(q)
    t = qq
    f = ONE
    for bit in bin(BLS_X)[3:]:
        val, t = _line(t, t, pp)
        f = mul(mul(f, f), val)
        if bit == "1":
            val, t = _line(t, qq, pp)
            f = mul(f, val)
    # conjugation = p^6 Frobenius; for the == 1 test f and f^-1 are equivalent
    return f


FINAL_EXP = (P**12 - 1) // R


def final_exp(f):
    return fpow(f, FINAL_EXP)


def pairing(p, q):
    return final_exp(miller(p, q))


def pairing_product_is_one(pairs) -> bool:
    f = ONE
    for p, q in pairs:
        f = mul(f, miller(p, q))
    return final_exp(f) == ONE
