"""BLS12-381 big-integer arithmetic for the test oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the shipped
package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may use it, and only as the checker.

The reference (lambdaclass/lambdaworks_kzg) keeps all of this arithmetic in the
un-vendored git dependency `lambdaworks-math` (Cargo.toml:15-16, no rev pin);
this file restates the published mathematics of BLS12-381 (constants:
SURVEY.md App. C) and the reference's own point codecs
(src/compression.rs:22-139).
"""
from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000  # |x|; the BLS parameter is -BLS_X

G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1

# ---------------------------------------------------------------- Fp helpers


def fp_inv(a: int) -> int:
    return pow(a, P - 2, P)


def fp_sqrt(a: int):
    """Square root in Fp (p = 3 mod 4).  Returns a root or None."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


def fr_inv(a: int) -> int:
    return pow(a, R - 2, R)


# ---------------------------------------------------------------- G1 (Jacobian)
# A point is None (infinity) or an affine tuple (x, y).  Internally scalar
# multiplication and sums use Jacobian triples for speed.

INF = None
G1 = (G1_X, G1_Y)


def g1_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def _jac_dbl(X1, Y1, Z1):
    if Z1 == 0 or Y1 == 0:
        return (1, 1, 0)
    A = X1 * X1 % P
    B = Y1 * Y1 % P
    C = B * B % P
    D = 2 * ((X1 + B) * (X1 + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y1 * Z1 % P
    return (X3, Y3, Z3)


def _jac_add(p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return _jac_dbl(X1, Y1, Z1)
        return (1, 1, 0)
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def _to_jac(pt):
    return (1, 1, 0) if pt is None else (pt[0], pt[1], 1)


def _from_jac(j):
    X, Y, Z = j
    if Z == 0:
        return None
    zi = fp_inv(Z)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_add(a, b):
    return _from_jac(_jac_add(_to_jac(a), _to_jac(b)))


def g1_neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def g1_mul(pt, k: int):
    """[k]pt for any non-negative integer k (no reduction mod r: callers that
    test subgroup membership rely on that)."""
    if pt is None or k == 0:
        return None
    acc = (1, 1, 0)
    base = _to_jac(pt)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(*acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return _from_jac(acc)


def g1_sum(points):
    acc = (1, 1, 0)
    for p in points:
        acc = _jac_add(acc, _to_jac(p))
    return _from_jac(acc)


def g1_msm(points, scalars, c: int = 8):
    """Generic bucket MSM (unsigned c-bit windows).  Any correct MSM yields the
    same group element, so the window size is a free choice here."""
    assert len(points) == len(scalars)
    if not points:
        return None
    nbits = max(1, max(s.bit_length() for s in scalars))
    nwin = (nbits + c - 1) // c
    jp = [_to_jac(p) for p in points]
    total = (1, 1, 0)
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = _jac_dbl(*total)
        buckets = [(1, 1, 0)] * (1 << c)
        for p, s in zip(jp, scalars):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = _jac_add(buckets[d], p)
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _from_jac(total)


def g1_in_subgroup(pt) -> bool:
    """src/compression.rs:22-27: [r]P == O."""
    return g1_mul(pt, R) is None


# ---------------------------------------------------------------- G1 codecs


class PointError(ValueError):
    pass


def g1_compress(pt) -> bytes:
    """src/compression.rs:33-60."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if (P - y) % P < y:
        b[0] |= 0x20
    return bytes(b)


def g1_decompress(data: bytes, *, strict: bool = False):
    """src/compression.rs:62-103 (strict=False).

    strict=True applies the c-kzg/blst (ZCash) rules used by the YAML vectors:
    infinity must be exactly c0 00.., x must be canonical (< p).
    """
    if len(data) != 48:
        raise PointError("length")
    b0 = data[0]
    if not b0 & 0x80:
        raise PointError("not compressed")
    if b0 & 0x40:
        if strict and (b0 != 0xC0 or any(data[1:])):
            raise PointError("bad infinity")
        return None
    x = int.from_bytes(bytes([b0 & 0x1F]) + data[1:], "big")
    if strict and x >= P:
        raise PointError("x >= p")
    x %= P
    y = fp_sqrt(x * x * x + 4)
    if y is None:
        raise PointError("not on curve")
    lo, hi = (y, P - y) if y < P - y else (P - y, y)
    y = hi if b0 & 0x20 else lo
    pt = (x, y)
    if not g1_in_subgroup(pt):
        raise PointError("not in subgroup")
    return pt


# ---------------------------------------------------------------- Fp2 / G2

def fp2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def fp2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fp2_sqr(a):
    return fp2_mul(a, a)


def fp2_inv(a):
    n = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * n % P, (-a[1]) * n % P)


def fp2_scalar(a, k):
    return (a[0] * k % P, a[1] * k % P)


def fp2_pow(a, e: int):
    r = (1, 0)
    for bit in bin(e)[2:]:
        r = fp2_sqr(r)
        if bit == "1":
            r = fp2_mul(r, a)
    return r


def fp2_sqrt(a):
    """A square root in Fp2 = Fp[u]/(u^2+1), or None.  (complex method)"""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-a0) % P)
        return (0, s) if s is not None else None
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    half = fp_inv(2)
    for nn in (n, P - n):
        t = (a0 + nn) * half % P
        x0 = fp_sqrt(t)
        if x0 is None or x0 == 0:
            continue
        x1 = a1 * fp_inv(2 * x0) % P
        if fp2_sqr((x0, x1)) == (a0, a1):
            return (x0, x1)
    return None


B2 = (4, 4)  # twist: y^2 = x^3 + 4(1+u)


def g2_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return fp2_sub(fp2_sqr(y), fp2_add(fp2_mul(fp2_sqr(x), x), B2)) == (0, 0)


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if y1 != y2 or y1 == (0, 0):
            return None
        lam = fp2_mul(fp2_scalar(fp2_sqr(x1), 3), fp2_inv(fp2_scalar(y1, 2)))
    else:
        lam = fp2_mul(fp2_sub(y2, y1), fp2_inv(fp2_sub(x2, x1)))
    x3 = fp2_sub(fp2_sub(fp2_sqr(lam), x1), x2)
    y3 = fp2_sub(fp2_mul(lam, fp2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_neg(a):
    return None if a is None else (a[0], fp2_neg(a[1]))


def g2_mul(pt, k: int):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


def _fp2_lex_larger(y) -> bool:
    """ZCash sign rule for Fp2: compare the imaginary part first."""
    ny = fp2_neg(y)
    if y[1] != ny[1]:
        return y[1] > ny[1]
    return y[0] > ny[0]


def g2_decompress(data: bytes):
    """96-byte ZCash-format G2 point: bytes[0..48] = x.c1, bytes[48..96] = x.c0
    (src/compression.rs:105-139).  The reference ignores the sign bit
    (SURVEY App. A.10); we honour it, which is identical for the shipped
    setups (sign bits clear, and the verify booleans are invariant under a
    simultaneous negation of g2[0], g2[1])."""
    if len(data) != 96:
        raise PointError("length")
    b0 = data[0]
    if not b0 & 0x80:
        raise PointError("not compressed")
    if b0 & 0x40:
        return None
    x1 = int.from_bytes(bytes([b0 & 0x1F]) + data[1:48], "big") % P
    x0 = int.from_bytes(data[48:96], "big") % P
    x = (x0, x1)
    y = fp2_sqrt(fp2_add(fp2_mul(fp2_sqr(x), x), B2))
    if y is None:
        raise PointError("not on curve")
    if _fp2_lex_larger(y) != bool(b0 & 0x20):
        y = fp2_neg(y)
    return (x, y)


def g2_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80
    if _fp2_lex_larger(y):
        b[0] |= 0x20
    return bytes(b)
