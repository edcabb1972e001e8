"""CPU tier: the device math headers (compiled for the host with the PTX
primitives emulated) against the Python big-int oracle."""
import ctypes
import hashlib
import random

import pytest

from oracle.py import bls, kzg
from tests import _emul
from tests._emul import aff_bytes, aff_from, from_u32, u32

P, R = bls.P, bls.R
RP, RR = 1 << 384, 1 << 256


@pytest.fixture(scope="module")
def lib():
    return _emul.load()


EDGE_P = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << 380), (1 << 380) - 1]
EDGE_R = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, (1 << 254), (1 << 254) - 1]


def test_fp_ops(lib):
    rnd = random.Random(1)
    vals = EDGE_P + [rnd.randrange(P) for _ in range(200)]
    out = (ctypes.c_uint32 * 12)()
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        lib.emul_fp_mul(out, u32(a, 12), u32(b, 12))
        assert from_u32(out) == a * b * pow(RP, -1, P) % P
        lib.emul_fp_add(out, u32(a, 12), u32(b, 12))
        assert from_u32(out) == (a + b) % P
        lib.emul_fp_sub(out, u32(a, 12), u32(b, 12))
        assert from_u32(out) == (a - b) % P
        lib.emul_fp_neg(out, u32(a, 12))
        assert from_u32(out) == (-a) % P
        lib.emul_fp_to_mont(out, u32(a, 12))
        assert from_u32(out) == a * RP % P
        lib.emul_fp_from_mont(out, u32(a, 12))
        assert from_u32(out) == a * pow(RP, -1, P) % P


def test_fr_ops(lib):
    rnd = random.Random(2)
    vals = EDGE_R + [rnd.randrange(R) for _ in range(200)]
    out = (ctypes.c_uint32 * 8)()
    for i, a in enumerate(vals):
        b = vals[(i * 5 + 1) % len(vals)]
        lib.emul_fr_mul(out, u32(a, 8), u32(b, 8))
        assert from_u32(out) == a * b * pow(RR, -1, R) % R
        lib.emul_fr_add(out, u32(a, 8), u32(b, 8))
        assert from_u32(out) == (a + b) % R
        lib.emul_fr_sub(out, u32(a, 8), u32(b, 8))
        assert from_u32(out) == (a - b) % R


def test_inversions(lib):
    rnd = random.Random(3)
    out = (ctypes.c_uint32 * 12)()
    for a in [1, 2, P - 1, rnd.randrange(P), rnd.randrange(P)]:
        lib.emul_fp_inv(out, u32(a * RP % P, 12))            # binary GCD (fpinv.cuh): what the kernels call
        assert from_u32(out) == pow(a, -1, P) * RP % P
        lib.emul_fp_inv_fermat(out, u32(a * RP % P, 12))     # a^(p-2): the independent cross-check
        assert from_u32(out) == pow(a, -1, P) * RP % P
    out8 = (ctypes.c_uint32 * 8)()
    for a in [1, 2, R - 1, rnd.randrange(R)]:
        lib.emul_fr_inv(out8, u32(a * RR % R, 8))
        assert from_u32(out8) == pow(a, -1, R) * RR % R


def test_fr_from_be(lib):
    rnd = random.Random(4)
    out = (ctypes.c_uint32 * 8)()
    cases = [0, 1, R - 1, R, R + 1, 2 * R - 1, 2 * R, 2 * R + 5, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(100)]
    for v in cases:
        b = v.to_bytes(32, "big")
        lib.emul_fr_from_be32(out, b)
        assert from_u32(out) == v % R
        lib.emul_fr_from_be_words(out, b)
        assert from_u32(out) == v % R


def _rand_pt(rnd):
    return bls.g1_mul(bls.G1, rnd.randrange(1, R))


def test_g1_ops(lib):
    rnd = random.Random(5)
    out = (ctypes.c_uint8 * 96)()
    pts = [None, bls.G1, bls.g1_neg(bls.G1)] + [_rand_pt(rnd) for _ in range(6)]
    for a in pts:
        for b in pts + [a, bls.g1_neg(a)]:
            want = bls.g1_add(a, b)
            for op in (0, 1, 3):
                lib.emul_g1_op(out, aff_bytes(a), aff_bytes(b), op)
                assert aff_from(out) == want, (op, a is None, b is None)
        lib.emul_g1_op(out, aff_bytes(a), aff_bytes(a), 2)
        assert aff_from(out) == bls.g1_add(a, a)


def test_g1_scalar_mul(lib):
    rnd = random.Random(6)
    out = (ctypes.c_uint8 * 96)()
    p = _rand_pt(rnd)
    for k in [0, 1, 2, 3, R - 1, R, rnd.randrange(R), rnd.randrange(1 << 256)]:
        lib.emul_g1_mul(out, aff_bytes(p), u32(k, 8))
        assert aff_from(out) == bls.g1_mul(p, k)


def test_g1_scalar_mul_glv(lib):
    """k P = (k mod x^2) P + (k div x^2) (beta x, -y) for every canonical k < r (batched verification's multiples)."""
    rnd = random.Random(66)
    out = (ctypes.c_uint8 * 96)()
    x2 = bls.BLS_X ** 2
    ks = [0, 1, 2, x2 - 1, x2, x2 + 1, 2 * x2, R - 1, R - 2, (R - 1) // 2, (1 << 128) - 1, 1 << 128, 1 << 254]
    ks += [rnd.randrange(R) for _ in range(12)]
    for i, k in enumerate(ks):
        p = _rand_pt(rnd) if i % 4 == 0 else p
        lib.emul_g1_mul_glv(out, aff_bytes(p), u32(k % R, 8))
        assert aff_from(out) == bls.g1_mul(p, k % R), hex(k)
    lib.emul_g1_mul_glv(out, aff_bytes(None), u32(12345, 8))
    assert aff_from(out) is None


def _curve_point_not_in_subgroup(rnd):
    while True:
        x = rnd.randrange(P)
        y = bls.fp_sqrt(x * x * x + 4)
        if y is not None:
            pt = (x, y)
            if not bls.g1_in_subgroup(pt):
                return pt


def test_subgroup_check(lib):
    rnd = random.Random(7)
    for _ in range(4):
        assert lib.emul_g1_in_subgroup(aff_bytes(_rand_pt(rnd))) == 1
    assert lib.emul_g1_in_subgroup(aff_bytes(None)) == 1
    for _ in range(8):
        pt = _curve_point_not_in_subgroup(rnd)
        assert lib.emul_g1_on_curve(aff_bytes(pt)) == 1
        assert lib.emul_g1_in_subgroup(aff_bytes(pt)) == 0
    # small-order points: cofactor h = (x-1)^2/3; [r*h/3 ...] -- take h-torsion
    h = (bls.BLS_X + 1) ** 2 // 3  # (x-1)^2/3 with x negative
    pt = _curve_point_not_in_subgroup(rnd)
    tors = bls.g1_mul(pt, R)  # kills the G1 component, leaves the cofactor part
    assert tors is not None and bls.g1_mul(tors, h) is None
    assert lib.emul_g1_in_subgroup(aff_bytes(tors)) == 0
    assert lib.emul_g1_on_curve(aff_bytes((0, 2))) == 1  # compression.rs:160-165
    assert lib.emul_g1_in_subgroup(aff_bytes((0, 2))) == 0


def test_codec(lib):
    rnd = random.Random(8)
    out48 = (ctypes.c_uint8 * 48)()
    out96 = (ctypes.c_uint8 * 96)()
    for pt in [None, bls.G1, bls.g1_neg(bls.G1)] + [_rand_pt(rnd) for _ in range(5)]:
        lib.emul_g1_compress(out48, aff_bytes(pt))
        assert bytes(out48) == bls.g1_compress(pt)
        assert lib.emul_g1_decompress(out96, bytes(out48)) == 1
        assert aff_from(out96) == pt
    # rejection / laxness parity with the Python restatement of compression.rs
    bad = [bytes(48), bytes([0x40]) + bytes(47), bytes([0x80]) + bytes(47), bytes([0xE0]) + b"\x01" * 47]
    x_off = next(x for x in range(2, 50) if bls.fp_sqrt(x**3 + 4) is None)
    bad.append(bytes([0x80]) + x_off.to_bytes(47, "big"))
    ns = _curve_point_not_in_subgroup(rnd)
    bad.append(bls.g1_compress(ns))
    xg = bls.G1_X + P  # non-canonical x >= p (accepted & reduced by the reference)
    if xg < (1 << 381):
        b = bytearray(xg.to_bytes(48, "big")); b[0] |= 0x80; bad.append(bytes(b))
    for b in bad:
        try:
            want = bls.g1_decompress(b)
            ok = 1
        except bls.PointError:
            want, ok = None, 0
        got = lib.emul_g1_decompress(out96, b)
        assert got == ok, b.hex()
        if ok:
            assert aff_from(out96) == want


def test_sha256(lib):
    rnd = random.Random(9)
    out = (ctypes.c_uint8 * 32)()
    for ln in [0, 1, 55, 56, 63, 64, 65, 119, 120, 127, 128, 1000, 131152]:
        msg = bytes(rnd.randrange(256) for _ in range(ln))
        lib.emul_sha256(out, msg, ctypes.c_size_t(ln))
        assert bytes(out) == hashlib.sha256(msg).digest()


@pytest.mark.parametrize("c", [4, 5, 7, 8, 10, 12, 13, 14, 15, 16])
def test_glv_recode(lib, c):
    """csrc/recode.cuh: k = m + q x^2 with m, q < x^2; signed digits with an unsigned top window; every digit
    addresses an entry the table holds (glv_window_count)."""
    rnd = random.Random(10 + c)
    X2 = 0xD201000000010000 ** 2
    W = lib.emul_glv_windows(c)
    assert W == -(-128 // c)
    lib.emul_glv_window_count.restype = ctypes.c_uint32
    counts = [lib.emul_glv_window_count(c, j) for j in range(W)]
    assert counts[:-1] == [1 << (c - 1)] * (W - 1) and counts[-1] == ((X2 - 1) >> (c * (W - 1))) + 1
    digits = (ctypes.c_int32 * (2 * W))()
    q4, m4 = (ctypes.c_uint32 * 4)(), (ctypes.c_uint32 * 4)()
    special = [0, 1, R - 1, R - 2, (1 << 254) - 1, X2 - 1, X2, X2 + 1, (X2 - 2) * X2 + X2 - 1, 5 * X2 - 1, 5 * X2,
               ((1 << 128) - 1) % R, (X2 - 1) * X2]
    for k in special + [rnd.randrange(R) for _ in range(300)]:
        assert k < R
        lib.emul_glv_recode(q4, m4, digits, u32(k, 8), c)
        q, m = from_u32(q4), from_u32(m4)
        assert (q, m) == divmod(k, X2)
        d = list(digits)
        for h, v in ((0, m), (1, q)):
            dd = d[h * W:(h + 1) * W]
            assert sum(x << (c * j) for j, x in enumerate(dd)) == v
            assert all(-(1 << (c - 1)) < x <= (1 << (c - 1)) for x in dd[:-1])
            assert 0 <= dd[-1] <= counts[-1]
            assert all(abs(x) <= counts[j] for j, x in enumerate(dd))


@pytest.mark.parametrize("chunks", [1, 4, 32, 128])
def test_poly_eval_quot(lib, chunks):
    rnd = random.Random(11)
    n = 4096
    for trial in range(2):
        coeffs = [rnd.randrange(R) for _ in range(n)]
        if trial == 1:
            coeffs[100:] = [0] * (n - 100)
        z = rnd.randrange(R)
        arr = (ctypes.c_uint32 * (8 * n))()
        for i, cf in enumerate(coeffs):
            for j in range(8):
                arr[8 * i + j] = (cf >> (32 * j)) & 0xFFFFFFFF
        y8 = (ctypes.c_uint32 * 8)()
        q = (ctypes.c_uint32 * (8 * n))()
        lib.emul_poly_eval_quot(y8, q, arr, n, u32(z, 8), chunks)
        assert from_u32(y8) == kzg.horner(coeffs, z)
        want = kzg.ruffini(coeffs, z) + [0]
        got = [from_u32(q[8 * i : 8 * i + 8]) for i in range(n)]
        assert got == want


# ------------------------------------------------------------------ pairing
def _g2_bytes(q):
    (x0, x1), (y0, y1) = q
    return b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1))


def _tower_to_flat(raw576):
    """emul Fp12 memory order: c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (each Fp2 = c0, c1)
    -> coefficients of the flat basis of oracle/py/pairing.py."""
    vals = [int.from_bytes(raw576[48 * i : 48 * i + 48], "big") for i in range(12)]
    wpow = [0, 2, 4, 1, 3, 5]
    flat = [0] * 12
    for slot, i in enumerate(wpow):
        a0, a1 = vals[2 * slot], vals[2 * slot + 1]
        flat[i] = (flat[i] + a0 - a1) % P
        flat[i + 6] = (flat[i + 6] + a1) % P
    return tuple(flat)


def test_hard_part_identity():
    x = -bls.BLS_X
    assert 3 * ((P**4 - P**2 + 1) // R) == (x - 1) ** 2 * (x + P) * (x * x + P * P - 1) + 3


def test_g2_decompress(lib, setup_text):
    lines = setup_text.splitlines()
    out = (ctypes.c_uint8 * 192)()
    for ln in lines[2 + 4096 : 2 + 4096 + 4]:
        raw = bytes.fromhex(ln)
        assert lib.emul_g2_decompress(out, raw) == 1
        q = bls.g2_decompress(raw)
        assert bytes(out) == _g2_bytes(q)
        # flipped sign bit -> negated y
        raw2 = bytes([raw[0] ^ 0x20]) + raw[1:]
        assert lib.emul_g2_decompress(out, raw2) == 1
        assert bytes(out) == _g2_bytes(bls.g2_neg(q))
    assert lib.emul_g2_decompress(out, bytes([0xC0]) + bytes(95)) == 2
    assert lib.emul_g2_decompress(out, bytes(96)) == 0


def test_pairing_matches_oracle(lib, py_setup):
    from oracle.py import pairing as pr

    rnd = random.Random(12)
    out = (ctypes.c_uint8 * 576)()
    q = py_setup.g2[0]
    for k in (1, rnd.randrange(R)):
        p = bls.g1_mul(bls.G1, k)
        lib.emul_pairing_gt(out, aff_bytes(p), _g2_bytes(q))
        got = _tower_to_flat(bytes(out))
        e = pr.pairing(p, q)
        # device: conj(f)^(3 (p^12-1)/r) == e^-3
        assert pr.mul(got, pr.fpow(e, 3)) == pr.ONE


def test_pairing_check(lib, py_setup):
    rnd = random.Random(13)
    g2_0, g2_1 = py_setup.g2[0], py_setup.g2[1]
    tau = py_setup.tau
    for _ in range(2):
        a = rnd.randrange(1, R)
        A = bls.g1_mul(bls.G1, a)
        tA = bls.g1_mul(A, tau)
        # e(tau A, g2_0) * e(-A, g2_1) == 1
        assert lib.emul_pairing_check(aff_bytes(tA), _g2_bytes(g2_0), aff_bytes(bls.g1_neg(A)), _g2_bytes(g2_1)) == 1
        assert lib.emul_pairing_check(aff_bytes(tA), _g2_bytes(g2_0), aff_bytes(A), _g2_bytes(g2_1)) == 0
        wrong = bls.g1_mul(A, tau + 1)
        assert lib.emul_pairing_check(aff_bytes(wrong), _g2_bytes(g2_0), aff_bytes(bls.g1_neg(A)), _g2_bytes(g2_1)) == 0
    # infinity pairs contribute 1
    assert lib.emul_pairing_check(aff_bytes(None), _g2_bytes(g2_0), aff_bytes(None), _g2_bytes(g2_1)) == 1
    assert lib.emul_pairing_check(aff_bytes(bls.G1), _g2_bytes(g2_0), aff_bytes(None), _g2_bytes(g2_1)) == 0



def test_cyclotomic_squaring(lib, py_setup):
    p = bls.g1_mul(bls.G1, 424242)
    assert lib.emul_cyclotomic_sqr_check(aff_bytes(p), _g2_bytes(py_setup.g2[1])) == 1


def test_warp_cooperative_pairing(lib, py_setup):
    """pairing_warp.cuh (lanes emulated one after the other) == single-thread pairing, bit for bit."""
    g2_0, g2_1 = py_setup.g2[0], py_setup.g2[1]
    tau = py_setup.tau
    A = bls.g1_mul(bls.G1, 987654321)
    tA = bls.g1_mul(A, tau)
    one = ctypes.c_int(-1)
    assert lib.emul_warp_pairing_check(aff_bytes(tA), _g2_bytes(g2_0), aff_bytes(bls.g1_neg(A)), _g2_bytes(g2_1), ctypes.byref(one)) == 1
    assert one.value == 1
    assert lib.emul_warp_pairing_check(aff_bytes(tA), _g2_bytes(g2_0), aff_bytes(A), _g2_bytes(g2_1), ctypes.byref(one)) == 1
    assert one.value == 0
    assert lib.emul_warp_pairing_check(aff_bytes(None), _g2_bytes(g2_0), aff_bytes(bls.g1_neg(A)), _g2_bytes(g2_1), ctypes.byref(one)) == 1
    assert one.value == 0


def test_dedicated_squaring(lib):
    """mont_sqr (symmetric product phase + injected reduction) == a*a/R for both fields."""
    rnd = random.Random(99)
    out = (ctypes.c_uint32 * 12)()
    for a in EDGE_P + [rnd.randrange(P) for _ in range(400)]:
        lib.emul_fp_sqr(out, u32(a, 12))
        assert from_u32(out) == a * a * pow(RP, -1, P) % P, hex(a)
    out8 = (ctypes.c_uint32 * 8)()
    for a in EDGE_R + [rnd.randrange(R) for _ in range(400)]:
        lib.emul_fr_sqr(out8, u32(a, 8))
        assert from_u32(out8) == a * a * pow(RR, -1, R) % R, hex(a)


def test_karatsuba_multiplier(lib):
    rnd = random.Random(101)
    out = (ctypes.c_uint32 * 12)()
    vals = EDGE_P + [(1 << 192) - 1, ((1 << 192) - 1) << 189 | ((1 << 189) - 1), P - (1 << 192)] + [rnd.randrange(P) for _ in range(400)]
    vals = [v % P for v in vals]
    for i, a in enumerate(vals):
        b = vals[(i * 11 + 5) % len(vals)]
        lib.emul_fp_mul_k(out, u32(a, 12), u32(b, 12))
        assert from_u32(out) == a * b * pow(RP, -1, P) % P, (hex(a), hex(b))
    out8 = (ctypes.c_uint32 * 8)()
    rv = EDGE_R + [(1 << 128) - 1, R - (1 << 128)] + [rnd.randrange(R) for _ in range(300)]
    for i, a in enumerate(rv):
        b = rv[(i * 7 + 2) % len(rv)]
        lib.emul_fr_mul_k(out8, u32(a, 8), u32(b, 8))
        assert from_u32(out8) == a * b * pow(RR, -1, R) % R


def test_fp64_pipe_multiplier(lib):
    """csrc/fpdp.cuh (DFMA limb products, radix 2^48) == a b R^-1 mod p, same bits as the IMAD multiplier."""
    rnd = random.Random(202)
    out = (ctypes.c_uint32 * 12)()
    ref = (ctypes.c_uint32 * 12)()
    m48 = (1 << 48) - 1
    special = [sum(m48 << (48 * i) for i in range(8)) % P, sum((1 << 47) << (48 * i) for i in range(8)) % P,
               P - 1, P - 2, 1, 0, 2, (1 << 380), (1 << 381) - 1, m48, m48 << 48, P >> 1]
    vals = [v % P for v in EDGE_P + special] + [rnd.randrange(P) for _ in range(1500)]
    for i, a in enumerate(vals):
        b = vals[(i * 13 + 7) % len(vals)]
        lib.emul_fp_mul_dp(out, u32(a, 12), u32(b, 12))
        assert from_u32(out) == a * b * pow(RP, -1, P) % P, (hex(a), hex(b))
        lib.emul_fp_mul(ref, u32(a, 12), u32(b, 12))
        assert list(out) == list(ref)
        lib.emul_fp_sqr_dp(out, u32(a, 12))
        assert from_u32(out) == a * a * pow(RP, -1, P) % P, hex(a)


def test_gcd_inversion(lib):
    """csrc/fpinv.cuh: approximate binary GCD inversion == Fermat inversion == pow(x, -1, p) (Montgomery in/out)."""
    rnd = random.Random(303)
    out = (ctypes.c_uint32 * 12)()
    vals = [0, 1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 380, (1 << 381) - 1, (1 << 64) - 1, 1 << 32,
            (1 << 62) - 1, 1 << 61, 1 << 62, 1 << 63, (1 << 381) - (1 << 190), RP % P, pow(RP, 2, P)]
    vals += [rnd.randrange(1, P) for _ in range(3000)]
    vals += [rnd.randrange(1, 1 << rnd.randrange(1, 381)) for _ in range(1500)]
    vals += [P - rnd.randrange(1, 1 << rnd.randrange(1, 380)) for _ in range(1500)]
    for a in vals:
        a %= P
        lib.emul_fp_inv_gcd(out, u32(a, 12))   # a is taken as the Montgomery representation of a / R
        want = 0 if a == 0 else pow(a * pow(RP, -1, P) % P, -1, P) * RP % P
        assert from_u32(out) == want, hex(a)
