"""GPU tier: the CUDA path, called through the C ABI (ctypes), against the
oracle on the same inputs.  Bit-exact: every output is integer/byte data.

Re-expresses the reference's own tests (tests/lib_test.rs, src/compression.rs
unit tests) and adds the oracle-differential cases SURVEY.md §4 derives.
"""
import ctypes
import json
import os
import random

import pytest

pytestmark = pytest.mark.gpu

from oracle.py import bls, kzg  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
R = bls.R
GEN_HEX = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"


@pytest.fixture(scope="module")
def lw():
    import lambdaworks_kzg_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def settings8(lw):
    """Settings with an 8-bit fixed-base window (1.6 GB table: quick to build)."""
    lw.set_option("window_bits", 8)
    s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    yield s
    s.free()


@pytest.fixture(scope="module")
def settings13(lw):
    """The shipping configuration: 13-bit window (32 GiB table)."""
    lw.set_option("window_bits", 13)
    s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    yield s
    s.free()


@pytest.fixture(scope="module")
def ref(py_setup):
    return kzg.RefMode(py_setup)


def blob_from_coeffs(coeffs):
    b = b"".join((c % (1 << 256)).to_bytes(32, "big") for c in coeffs)
    return b + bytes(kzg.BYTES_PER_BLOB - len(b))


def edge_blobs():
    rnd = random.Random(4844)
    out = {
        "zero": bytes(kzg.BYTES_PER_BLOB),
        "one": blob_from_coeffs([1]),
        "x": blob_from_coeffs([0, 1]),
        "top_only": blob_from_coeffs([0] * 4095 + [7]),
        "all_ff": b"\xff" * kzg.BYTES_PER_BLOB,               # every word >= r (reduced mod r, App. A.1)
        "r_minus_1": blob_from_coeffs([R - 1] * 4096),
        "exactly_r": blob_from_coeffs([R] * 4096),              # == all-zero polynomial
        "r_plus_1": blob_from_coeffs([R + 1, 2 * R + 3, 2 * R - 1] + [0] * 10 + [R]),
        "random_full": bytes(rnd.randrange(256) for _ in range(kzg.BYTES_PER_BLOB)),
        "sparse": blob_from_coeffs([rnd.randrange(R) if i % 97 == 0 else 0 for i in range(4096)]),
        "repeat_digit": blob_from_coeffs([0x0101010101010101010101010101010101010101010101010101010101010101 % R] * 4096),
    }
    return out


# ------------------------------------------------------------------ setup
def test_setup_layout(lw, settings8, py_setup):
    g1 = settings8.g1_values_bytes()
    for i in (0, 1, 2, 4095):
        x, y = py_setup.g1[i]
        want = b"".join(int(v).to_bytes(8, "little") for v in _be_limbs(x)) + b"".join(int(v).to_bytes(8, "little") for v in _be_limbs(y)) + \
            b"".join(int(v).to_bytes(8, "little") for v in [0, 0, 0, 0, 0, 1])
        assert g1[144 * i: 144 * i + 144] == want, i
    g2 = settings8.g2_values_bytes()
    (x0, x1), (y0, y1) = py_setup.g2[1]
    want = b"".join(b"".join(int(v).to_bytes(8, "little") for v in _be_limbs(c)) for c in (x0, x1, y0, y1, 1, 0))
    assert g2[288:576] == want
    assert settings8.c.fs is not None


def _be_limbs(v):
    """u64 limbs, most significant first (src/srs.rs:131-153)."""
    return [(v >> (64 * (5 - k))) & 0xFFFFFFFFFFFFFFFF for k in range(6)]


def test_load_trusted_setup_bytes_and_badargs(lw, setup_text):
    lines = setup_text.splitlines()
    g1 = b"".join(bytes.fromhex(x) for x in lines[2: 2 + 4096])
    g2 = b"".join(bytes.fromhex(x) for x in lines[2 + 4096: 2 + 4096 + 65])
    with pytest.raises(lw.KzgError) as e:
        lw.load_trusted_setup(g1, g2, n1=4095, n2=65)
    assert e.value.code == lw.C_KZG_BADARGS  # lib.rs:716-718
    lw.set_option("window_bits", 8)
    s = lw.load_trusted_setup(g1, g2)
    try:
        assert lw.blob_to_kzg_commitment(blob_from_coeffs([0, 1]), s).hex() == lines[3]
    finally:
        s.free()
    bad = bytearray(g1)
    bad[48] ^= 0x01  # second point no longer on the curve / in the subgroup
    with pytest.raises(lw.KzgError) as e:
        lw.load_trusted_setup(bytes(bad), g2)
    assert e.value.code == lw.C_KZG_ERROR


def test_trusted_setup_4_parses_but_cannot_compute(lw):
    # src/srs.rs:282-295 only asserts that the short setup parses
    s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup_4.txt"))
    try:
        with pytest.raises(lw.KzgError) as e:
            lw.blob_to_kzg_commitment(bytes(kzg.BYTES_PER_BLOB), s)
        assert e.value.code == lw.C_KZG_ERROR
    finally:
        s.free()


# ------------------------------------------------------------------ reference tests re-expressed
def test_lib_test_simple_poly_1(lw, settings8):
    """tests/lib_test.rs:19-87: p = 1, z = 1 -> y = 1, proof = infinity, verifies."""
    blob = blob_from_coeffs([1])
    z = (1).to_bytes(32, "big")
    proof, y = lw.compute_kzg_proof(blob, z, settings8)
    assert int.from_bytes(y, "big") == 1
    assert proof == bytes([0xC0]) + bytes(47)
    commitment = bytes.fromhex(GEN_HEX)  # commit(1) = G for any SRS
    assert lw.verify_kzg_proof(commitment, z, y, proof, settings8) is True


def test_lib_test_simple_poly_2(lw, settings8, setup_text):
    """tests/lib_test.rs:89-167: p = X, z = 2 -> y = 2, proof = g1[0], commitment = g1[1]."""
    lines = setup_text.splitlines()
    blob = blob_from_coeffs([0, 1])
    z = (2).to_bytes(32, "big")
    proof, y = lw.compute_kzg_proof(blob, z, settings8)
    assert y == z
    assert proof.hex() == lines[2]
    commitment = lw.blob_to_kzg_commitment(blob, settings8)
    assert commitment.hex() == lines[3]
    assert lw.verify_kzg_proof(commitment, z, y, proof, settings8) is True


def test_lib_test_batch_proof(lw, settings8, ref):
    """tests/lib_test.rs:169-260: both pairs through verify_blob_kzg_proof_batch(n = 2).
    (The reference test feeds compute_kzg_proof outputs at z = 1, 2 -- valid for any
    Fiat-Shamir z because deg <= 1: the quotient is constant.)"""
    b1, b2 = blob_from_coeffs([1]), blob_from_coeffs([0, 1])
    p1, _ = lw.compute_kzg_proof(b1, (1).to_bytes(32, "big"), settings8)
    p2, _ = lw.compute_kzg_proof(b2, (2).to_bytes(32, "big"), settings8)
    c1, c2 = lw.blob_to_kzg_commitment(b1, settings8), lw.blob_to_kzg_commitment(b2, settings8)
    assert lw.verify_blob_kzg_proof_batch([b1, b2], [c1, c2], [p1, p2], settings8) is True
    assert ref.verify_blob_kzg_proof_batch([b1, b2], [c1, c2], [p1, p2]) is True


def test_lib_test_read_srs(lw, settings8):
    """tests/lib_test.rs:262-291: first SRS point compresses to the generator."""
    assert lw.blob_to_kzg_commitment(blob_from_coeffs([1]), settings8).hex() == GEN_HEX


def test_compression_kats(lw, settings8):
    """src/compression.rs:168-221 through the ABI: generator / 2G / infinity / KAT point round trips."""
    s = settings8
    two_g = bls.g1_compress(bls.g1_mul(bls.G1, 2))
    kat = bytes.fromhex("8d0c6eeadd3f8529d67246f77404a4ac2d9d7fd7d50cf103d3e6abb9003e5e36d8f322663ebced6707a7f46d97b7566d")
    for enc in (bytes.fromhex(GEN_HEX), two_g, bytes([0xC0]) + bytes(47), kat):
        # compute_blob_kzg_proof decodes (and validates) the commitment; for the zero
        # blob the proof is infinity whatever the commitment is
        assert lw.compute_blob_kzg_proof(bytes(kzg.BYTES_PER_BLOB), enc, s) == bytes([0xC0]) + bytes(47)
    for bad in (bytes(48), bytes([0x80]) + bytes(47), bls.g1_compress((0, 2))):  # (0,2): on curve, not in G1
        with pytest.raises(lw.KzgError) as e:
            lw.compute_blob_kzg_proof(bytes(kzg.BYTES_PER_BLOB), bad, s)
        assert e.value.code == lw.C_KZG_ERROR


# ------------------------------------------------------------------ oracle differential
def test_ref_mode_kats(lw, settings8):
    kats = json.load(open(os.path.join(GOLDEN, "ref_mode_kats.json")))
    idx = json.load(open(os.path.join(GOLDEN, "ckzg_le_vectors.json")))["blobs"]
    blobs_bin = open(os.path.join(GOLDEN, "ckzg_le_blobs.bin"), "rb").read()
    for k in kats:
        off, ln = idx[k["blob"]]
        blob = blobs_bin[off: off + ln]
        com = lw.blob_to_kzg_commitment(blob, settings8)
        assert com.hex() == k["commitment"], k["name"]
        proof, y = lw.compute_kzg_proof(blob, (2).to_bytes(32, "big"), settings8)
        assert (proof.hex(), y.hex()) == (k["z2_proof"], k["z2_y"]), k["name"]
        bp = lw.compute_blob_kzg_proof(blob, com, settings8)
        assert bp.hex() == k["blob_proof"], k["name"]
        assert lw.verify_blob_kzg_proof(blob, com, bp, settings8) is True
        assert lw.verify_kzg_proof(com, bytes.fromhex(k["fs_z"]), bytes.fromhex(k["fs_y"]), bp, settings8) is True


@pytest.mark.parametrize("which", ["settings8", "settings13"])
def test_edge_blobs_vs_oracle(lw, ref, request, which):
    s = request.getfixturevalue(which)
    blobs = edge_blobs()
    names = list(blobs)
    cat = b"".join(blobs[n] for n in names)
    coms, proofs, st = lw.commit_and_prove_batch(cat, len(names), s)
    assert st == [0] * len(names)
    for n, c, p in zip(names, coms, proofs):
        want_c = ref.blob_to_kzg_commitment(blobs[n])
        assert c == want_c, n
        assert p == ref.compute_blob_kzg_proof(blobs[n], want_c), n
        assert lw.blob_to_kzg_commitment(blobs[n], s) == want_c, n
    ok = lw.verify_blob_kzg_proof_batch([blobs[n] for n in names], coms, proofs, s)
    assert ok is True


def test_synthetic_batch_vs_oracle(lw, settings13, ref):
    n = 24
    blobs = [lw.synth_blob_host(k) for k in range(n)]
    coms, proofs, st = lw.commit_and_prove_batch(b"".join(blobs), n, settings13)
    assert st == [0] * n
    for k in range(n):
        c = ref.blob_to_kzg_commitment(blobs[k])
        assert coms[k] == c, k
        assert proofs[k] == ref.compute_blob_kzg_proof(blobs[k], c), k
    # the separate entry points give the same bytes
    coms2, st2 = lw.blob_to_kzg_commitment_batch(b"".join(blobs), n, settings13)
    proofs2, st3 = lw.compute_blob_kzg_proof_batch(b"".join(blobs), b"".join(coms), n, settings13)
    assert coms2 == coms and proofs2 == proofs and st2 == st3 == [0] * n
    zs = [(k * 7919 + 1).to_bytes(32, "big") for k in range(n)]
    pp, yy, st4 = lw.compute_kzg_proof_batch(b"".join(blobs), b"".join(zs), n, settings13)
    for k in range(0, n, 5):
        wp, wy = ref.compute_kzg_proof(blobs[k], zs[k])
        assert (pp[k], yy[k]) == (wp, wy)


def test_fuzz_corpus_differential(lw, settings8, ref):
    meta = json.load(open(os.path.join(GOLDEN, "fuzz_corpus.json")))
    raw = open(os.path.join(GOLDEN, "fuzz_corpus.bin"), "rb").read()
    B = kzg.BYTES_PER_BLOB

    def both(fn_gpu, fn_ref):
        try:
            want = ("ok", fn_ref())
        except kzg.KzgError as e:
            want = ("err", e.code)
        try:
            got = ("ok", fn_gpu())
        except lw.KzgError as e:
            got = ("err", e.code)
        assert got == want

    for case in meta["cases"]:
        off, ln = meta["blobs"][case["data"]]
        d = raw[off: off + ln]
        suite = case["suite"]
        if suite == "blob_to_kzg_commitment":
            both(lambda: lw.blob_to_kzg_commitment(d, settings8), lambda: ref.blob_to_kzg_commitment(d))
        elif suite == "compute_kzg_proof":
            both(lambda: lw.compute_kzg_proof(d[:B], d[B:], settings8), lambda: ref.compute_kzg_proof(d[:B], d[B:]))
        elif suite == "compute_blob_kzg_proof":
            both(lambda: lw.compute_blob_kzg_proof(d[:B], d[B:], settings8), lambda: ref.compute_blob_kzg_proof(d[:B], d[B:]))
        elif suite == "verify_kzg_proof":
            both(lambda: lw.verify_kzg_proof(d[:48], d[48:80], d[80:112], d[112:], settings8),
                 lambda: ref.verify_kzg_proof(d[:48], d[48:80], d[80:112], d[112:]))
        elif suite == "verify_blob_kzg_proof":
            both(lambda: lw.verify_blob_kzg_proof(d[:B], d[B:B + 48], d[B + 48:], settings8),
                 lambda: ref.verify_blob_kzg_proof(d[:B], d[B:B + 48], d[B + 48:]))
        elif suite == "verify_blob_kzg_proof_batch":
            n = case["n"]
            blobs = [d[i * B:(i + 1) * B] for i in range(n)]
            cs = [d[n * B + 48 * i: n * B + 48 * i + 48] for i in range(n)]
            ps = [d[n * B + 48 * n + 48 * i: n * B + 48 * n + 48 * i + 48] for i in range(n)]
            both(lambda: lw.verify_blob_kzg_proof_batch(blobs, cs, ps, settings8), lambda: ref.verify_blob_kzg_proof_batch(blobs, cs, ps))


# ------------------------------------------------------------------ verification outcomes
def test_verify_negative_and_errors(lw, settings8, ref):
    rnd = random.Random(99)
    blob = lw.synth_blob_host(12345)
    com = lw.blob_to_kzg_commitment(blob, settings8)
    proof = lw.compute_blob_kzg_proof(blob, com, settings8)
    assert lw.verify_blob_kzg_proof(blob, com, proof, settings8) is True
    g = bytes.fromhex(GEN_HEX)
    assert lw.verify_blob_kzg_proof(blob, com, g, settings8) is False
    assert lw.verify_blob_kzg_proof(blob, g, proof, settings8) is False
    blob2 = bytearray(blob); blob2[77] ^= 1
    assert lw.verify_blob_kzg_proof(bytes(blob2), com, proof, settings8) is False
    z = rnd.randrange(R).to_bytes(32, "big")
    p2, y2 = lw.compute_kzg_proof(blob, z, settings8)
    assert lw.verify_kzg_proof(com, z, y2, p2, settings8) is True
    ybad = ((int.from_bytes(y2, "big") + 1) % R).to_bytes(32, "big")
    assert lw.verify_kzg_proof(com, z, ybad, p2, settings8) is False
    # non-canonical y (y + r) is reduced, not rejected (App. A.1)
    ynon = int.from_bytes(y2, "big") + R
    if ynon < 1 << 256:
        assert lw.verify_kzg_proof(com, z, ynon.to_bytes(32, "big"), p2, settings8) is True
    for badpt in (bytes(48), bls.g1_compress((0, 2))):
        with pytest.raises(lw.KzgError):
            lw.verify_kzg_proof(badpt, z, y2, p2, settings8)
        with pytest.raises(lw.KzgError):
            lw.verify_blob_kzg_proof(blob, com, badpt, settings8)
    # infinity proof / commitment are legal encodings
    inf = bytes([0xC0]) + bytes(47)
    assert lw.verify_kzg_proof(inf, z, bytes(32), inf, settings8) is True  # p = 0
    assert lw.verify_kzg_proof(com, z, y2, inf, settings8) is ref.verify_kzg_proof(com, z, y2, inf)


def test_batch_verify_outcomes(lw, settings8, ref):
    n = 5
    blobs = [lw.synth_blob_host(1000 + k) for k in range(n)]
    coms, proofs, _ = lw.commit_and_prove_batch(b"".join(blobs), n, settings8)
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, proofs, settings8) is True
    assert lw.verify_blob_kzg_proof_batch([], [], [], settings8) is False          # lib.rs:538-543
    assert lw.verify_blob_kzg_proof_batch(blobs[:1], coms[:1], proofs[:1], settings8) is True
    bad = list(proofs); bad[3] = bytes.fromhex(GEN_HEX)
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, bad, settings8) is False
    assert ref.verify_blob_kzg_proof_batch(blobs, coms, bad) is False
    swapped = [proofs[1], proofs[0]] + proofs[2:]
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, swapped, settings8) is False
    inval = list(coms); inval[2] = bytes(48)
    with pytest.raises(lw.KzgError) as e:
        lw.verify_blob_kzg_proof_batch(blobs, inval, proofs, settings8)
    assert e.value.code == lw.C_KZG_ERROR
    # phases == monolithic call, for 1, 2 and 3 simulated ranks
    for world in (1, 2, 3):
        tuples, shards = [], []
        for r in range(world):
            first, cnt = lw.shard_range(n, world, r)
            shards.append((first, cnt))
        # single process: phase1 of a rank must be followed by its phase2, so gather tuples first
        for first, cnt in shards:
            tuples.append(lw.verify_batch_phase1(b"".join(blobs[first:first + cnt]), b"".join(coms[first:first + cnt]),
                                                 b"".join(proofs[first:first + cnt]), cnt, settings8))
        all_t = b"".join(tuples)
        parts = []
        for first, cnt in shards:
            lw.verify_batch_phase1(b"".join(blobs[first:first + cnt]), b"".join(coms[first:first + cnt]),
                                   b"".join(proofs[first:first + cnt]), cnt, settings8)
            parts.append(lw.verify_batch_phase2(all_t, n, first, cnt, settings8) if cnt else bytes(288))
        assert lw.verify_batch_phase3(b"".join(parts), world, settings8) is True


@pytest.mark.parametrize("n,super_blobs,chunk", [(67, 16384, 256), (130, 64, 256), (300, 128, 50)])
def test_batch_verify_staged_pipeline(lw, settings8, n, super_blobs, chunk):
    """Large-batch verification path: group SHA-256 kernel with a batch size that is not a multiple of its 8 blobs
    per warp, several staging super-batches with a partial last one, odd chunk sizes, and the batch challenge r
    absorbed chunk by chunk -- r is checked against hashlib over the tuples (utils.rs:166-206) and against the
    monolithic hash of the multi-GPU phase 2."""
    import hashlib

    blobs = [lw.synth_blob_host(5000 + k) for k in range(n)]
    flat = b"".join(blobs)
    coms, proofs, st = lw.commit_and_prove_batch(flat, n, settings8)
    assert st == [0] * n
    p2, st2 = lw.compute_blob_kzg_proof_batch(flat, b"".join(coms), n, settings8)   # hash not hidden under an MSM here
    assert st2 == [0] * n and p2 == proofs
    lw.set_option("verify_super_blobs", super_blobs)
    lw.set_option("chunk_blobs", chunk)
    try:
        assert lw.verify_blob_kzg_proof_batch(blobs, coms, proofs, settings8) is True
        r_chunked = lw.debug_batch_challenge(settings8)
        tuples = lw.verify_batch_phase1(flat, b"".join(coms), b"".join(proofs), n, settings8)
        assert len(tuples) == 160 * n and tuples[:48] == coms[0] and tuples[112:160] == proofs[0]
        msg = b"RCKZGBATCH___V1_" + (4096).to_bytes(8, "little") + n.to_bytes(8, "little") + tuples
        expect = int.from_bytes(hashlib.sha256(msg).digest(), "big") % R
        assert r_chunked == expect
        part = lw.verify_batch_phase2(tuples, n, 0, n, settings8)
        assert lw.debug_batch_challenge(settings8) == expect
        assert lw.verify_batch_phase3(part, 1, settings8) is True
        bad = list(proofs)
        bad[n - 1] = coms[0]
        assert lw.verify_blob_kzg_proof_batch(blobs, coms, bad, settings8) is False
        bad = list(proofs)
        bad[0], bad[1] = bad[1], bad[0]
        assert lw.verify_blob_kzg_proof_batch(blobs, coms, bad, settings8) is False
        inval = list(coms)
        inval[n // 2] = bytes(48)
        with pytest.raises(lw.KzgError):
            lw.verify_blob_kzg_proof_batch(blobs, inval, proofs, settings8)
    finally:
        lw.set_option("verify_super_blobs", 16384)
        lw.set_option("chunk_blobs", 256)


def test_g1_lincomb(lw, py_setup):
    rnd = random.Random(5)
    n = 37
    pts = [py_setup.g1[rnd.randrange(4096)] for _ in range(n)]
    pts[3] = None
    sc = [rnd.randrange(1 << 256) for _ in range(n)]
    sc[5] = 0
    pb = b"".join(bytes(96) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p in pts)
    sb = b"".join(s.to_bytes(32, "big") for s in sc)
    want = bls.g1_compress(bls.g1_msm(pts, [s % R for s in sc]))
    assert lw.g1_lincomb(pb, sb, n) == want
    assert lw.g1_lincomb(b"", b"", 0) == bytes([0xC0]) + bytes(47)


# ------------------------------------------------------------------ hand-built settings (fs == NULL) and bad SRS
def test_hand_built_settings(lw, settings8, py_setup):
    from lambdaworks_kzg_b200.api import CKZGSettings

    g1 = ctypes.create_string_buffer(settings8.g1_values_bytes(), 4096 * 144)
    g2 = ctypes.create_string_buffer(settings8.g2_values_bytes(), 65 * 288)
    s = CKZGSettings(None, ctypes.cast(g1, ctypes.c_void_p), ctypes.cast(g2, ctypes.c_void_p))
    lw.set_option("window_bits", 8)
    blob = lw.synth_blob_host(7)
    assert lw.blob_to_kzg_commitment(blob, s) == lw.blob_to_kzg_commitment(blob, settings8)
    # swap two SRS points in place: content hash changes -> new context, new result
    raw = bytearray(g1.raw)
    raw[0:144], raw[144:288] = raw[144:288], raw[0:144]
    ctypes.memmove(g1, bytes(raw), len(raw))
    swapped = kzg.Setup(g1=[py_setup.g1[1], py_setup.g1[0]] + py_setup.g1[2:], g2=py_setup.g2, tau=None)
    small = blob_from_coeffs([5, 9, 11])
    assert lw.blob_to_kzg_commitment(small, s) == kzg.RefMode(swapped, generic=True).blob_to_kzg_commitment(small)
    # an infinity / off-curve g1 value makes every call fail (SRS re-hydration, srs.rs:155-172)
    raw[288:288 + 96] = bytes(96)
    ctypes.memmove(g1, bytes(raw), len(raw))
    with pytest.raises(lw.KzgError) as e:
        lw.blob_to_kzg_commitment(small, s)
    assert e.value.code == lw.C_KZG_ERROR


# ------------------------------------------------------------------ batched-affine MSM kernel (large-batch path)
@pytest.fixture
def force_batch_affine(lw):
    """Route even tiny batches through msm_gather_ba_kernel (normally used from 512 blobs up)."""
    old = lw.get_option("msm_ba_min_blobs")
    lw.set_option("msm_algo", 1)
    lw.set_option("msm_ba_min_blobs", 1)
    yield
    lw.set_option("msm_ba_min_blobs", old)
    lw.set_option("msm_ba_variant", 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_batch_affine_edge_blobs_vs_oracle(lw, ref, settings8, force_batch_affine, variant):
    lw.set_option("msm_ba_variant", variant)
    blobs = edge_blobs()
    names = list(blobs)
    coms, proofs, st = lw.commit_and_prove_batch(b"".join(blobs[n] for n in names), len(names), settings8)
    assert st == [0] * len(names)
    for n, c, p in zip(names, coms, proofs):
        want_c = ref.blob_to_kzg_commitment(blobs[n])
        assert c == want_c, n
        assert p == ref.compute_blob_kzg_proof(blobs[n], want_c), n


def test_batch_affine_synthetic_vs_oracle(lw, ref, settings13, force_batch_affine):
    n = 12
    blobs = [lw.synth_blob_host(k) for k in range(n)]
    coms, proofs, st = lw.commit_and_prove_batch(b"".join(blobs), n, settings13)
    assert st == [0] * n
    for k in range(n):
        c = ref.blob_to_kzg_commitment(blobs[k])
        assert coms[k] == c, k
        assert proofs[k] == ref.compute_blob_kzg_proof(blobs[k], c), k


def test_batch_affine_repeated_srs_points(lw, py_setup, force_batch_affine):
    """Equal table entries meet in one accumulator slot (T == A: doubling; T == -A: cancellation): a hand-built
    SRS whose 4096 points are all the same point, blobs with few distinct words."""
    from lambdaworks_kzg_b200.api import CKZGSettings

    lw.set_option("window_bits", 8)
    tmp = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    raw = bytearray(tmp.g1_values_bytes())
    g2raw = tmp.g2_values_bytes()
    tmp.free()
    for i in range(4096):
        if i != 1:
            raw[144 * i: 144 * i + 144] = raw[144:288]      # every point = g1[1]
    g1 = ctypes.create_string_buffer(bytes(raw), 4096 * 144)
    g2 = ctypes.create_string_buffer(g2raw, 65 * 288)
    s = CKZGSettings(None, ctypes.cast(g1, ctypes.c_void_p), ctypes.cast(g2, ctypes.c_void_p))
    rnd = random.Random(77)
    w = [rnd.randrange(R) for _ in range(3)]
    cases = [
        [w[0]] * 4096,
        [w[0], R - w[0]] * 2048,                               # sums to zero: commitment = infinity
        [w[rnd.randrange(3)] for _ in range(4096)],
        [1] * 4096,
        [(1 << 7)] * 4096,                                     # digit 2^(c-1) in the lowest window
        [w[0]] * 64 + [0] * 4032,
    ]
    blobs = [blob_from_coeffs(c) for c in cases]
    coms, st = lw.blob_to_kzg_commitment_batch(b"".join(blobs), len(blobs), s)
    assert st == [0] * len(blobs)
    P1 = py_setup.g1[1]
    for c, got in zip(cases, coms):
        want = bls.g1_compress(bls.g1_mul(P1, sum(c) % R) if sum(c) % R else None)
        assert got == want
    lw.set_option("msm_algo", 0)
    coms0, _ = lw.blob_to_kzg_commitment_batch(b"".join(blobs), len(blobs), s)
    lw.set_option("msm_algo", 1)
    assert coms0 == coms


def test_large_batch_with_degenerate_blobs_both_kernels(lw, ref, settings13):
    """512 blobs, a handful of them degenerate: both MSM kernels, identical bytes, oracle on the odd ones.  (This
    shape crashed msm_finalize_kernel before the single-exit rewrite of the group law -- profiles/r01_sanitizer.md.)"""
    n = 512
    blobs = [lw.synth_blob_host(k) for k in range(n)]
    special = {1: bytes(kzg.BYTES_PER_BLOB), 2: blob_from_coeffs([123456789] * 4096), 3: blob_from_coeffs([0] * 4095 + [1]),
               4: blob_from_coeffs([R - 1] * 4096), 5: blob_from_coeffs([1] * 4096), 6: blob_from_coeffs([R + 5] * 4096), 300: blob_from_coeffs([7])}
    for k, b in special.items():
        blobs[k] = b
    cat = b"".join(blobs)
    out = {}
    for algo in (0, 1):
        lw.set_option("msm_algo", algo)
        coms, proofs, st = lw.commit_and_prove_batch(cat, n, settings13)
        assert st == [0] * n
        out[algo] = (coms, proofs)
    lw.set_option("msm_algo", 1)
    assert out[0] == out[1]
    for k in list(special) + [0, 511]:
        c = ref.blob_to_kzg_commitment(blobs[k])
        assert out[1][0][k] == c, k
        assert out[1][1][k] == ref.compute_blob_kzg_proof(blobs[k], c), k


@pytest.mark.parametrize("n", [256, 257, 300])
def test_batch_sizes_around_the_kernel_switch(lw, ref, settings8, n):
    """Chunks of 256 blobs go to the batched-affine kernel, the remainder chunk (1 .. 255 blobs) to the XYZZ kernel:
    same bytes as an XYZZ-only run, oracle on blobs either side of the boundary."""
    cat = b"".join(lw.synth_blob_host(k) for k in range(n))
    lw.set_option("msm_algo", 1)
    coms, proofs, st = lw.commit_and_prove_batch(cat, n, settings8)
    lw.set_option("msm_algo", 0)
    coms0, proofs0, st0 = lw.commit_and_prove_batch(cat, n, settings8)
    lw.set_option("msm_algo", 1)
    assert st == st0 == [0] * n
    assert coms == coms0 and proofs == proofs0
    for k in (0, 255, n - 1):
        blob = lw.synth_blob_host(k)
        c = ref.blob_to_kzg_commitment(blob)
        assert coms[k] == c and proofs[k] == ref.compute_blob_kzg_proof(blob, c), k


@pytest.mark.parametrize("n", [1, 130, 300, 700])
def test_device_api_odd_batch_sizes(lw, ref, settings8, n):
    """Device-pointer entry point with batch sizes that leave a remainder chunk (a small chunk needs more block
    partials per blob than a full one: the slots are sized for the whole chunk plan before anything is enqueued)."""
    import torch

    dev = torch.device("cuda", 0)
    blobs = torch.empty(n * kzg.BYTES_PER_BLOB, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lw.synth_blobs_device(blobs.data_ptr(), 0, n, st)
    coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    status = torch.ones(n, dtype=torch.int32, device=dev)
    lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, settings8, st, status.data_ptr())
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    cb, pb = bytes(coms.cpu().numpy().tobytes()), bytes(proofs.cpu().numpy().tobytes())
    hc, hp, hs = lw.commit_and_prove_batch(bytes(blobs.cpu().numpy().tobytes()), n, settings8)
    assert hs == [0] * n and b"".join(hc) == cb and b"".join(hp) == pb
    for k in sorted({0, n // 2, n - 1}):
        blob = lw.synth_blob_host(k)
        c = ref.blob_to_kzg_commitment(blob)
        assert cb[48 * k: 48 * k + 48] == c and pb[48 * k: 48 * k + 48] == ref.compute_blob_kzg_proof(blob, c), k


# ------------------------------------------------------------------ device API + full-size properties
def test_device_api_and_large_batch_properties(lw, settings13, ref):
    import torch

    n = 1024
    dev = torch.device("cuda", 0)
    blobs = torch.empty(n * kzg.BYTES_PER_BLOB, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(blobs.data_ptr(), 0, n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = bytes(blobs[: 2 * kzg.BYTES_PER_BLOB].cpu().numpy().tobytes())
    assert host[: kzg.BYTES_PER_BLOB] == lw.synth_blob_host(0) and host[kzg.BYTES_PER_BLOB:] == lw.synth_blob_host(1)
    coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    status = torch.ones(n, dtype=torch.int32, device=dev)
    lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, settings13,
                                     torch.cuda.current_stream().cuda_stream, status.data_ptr())
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    cb, pb = bytes(coms.cpu().numpy().tobytes()), bytes(proofs.cpu().numpy().tobytes())
    for k in (0, 1, 511, 512, 1023):
        blob = lw.synth_blob_host(k)
        c = ref.blob_to_kzg_commitment(blob)
        assert cb[48 * k: 48 * k + 48] == c, k
        assert pb[48 * k: 48 * k + 48] == ref.compute_blob_kzg_proof(blob, c), k
    # size-independent property at the full batch: every (blob, C, pi) verifies, and one
    # corrupted proof flips the batch result
    all_blobs = bytes(blobs.cpu().numpy().tobytes())
    bl = [all_blobs[i * kzg.BYTES_PER_BLOB:(i + 1) * kzg.BYTES_PER_BLOB] for i in range(n)]
    cl = [cb[48 * i: 48 * i + 48] for i in range(n)]
    pl = [pb[48 * i: 48 * i + 48] for i in range(n)]
    assert lw.verify_blob_kzg_proof_batch(bl, cl, pl, settings13) is True
    pl[777] = bytes.fromhex(GEN_HEX)
    assert lw.verify_blob_kzg_proof_batch(bl, cl, pl, settings13) is False
    # checksum of checksums: sum of commitments == commitment of the summed blob (linearity)
    total = [0] * 4096
    for k in range(8):
        for i in range(4096):
            total[i] += int.from_bytes(bl[k][32 * i: 32 * i + 32], "big")
    summed = blob_from_coeffs([t % R for t in total])
    want = bls.g1_compress(bls.g1_sum(bls.g1_decompress(cl[k]) for k in range(8)))
    assert lw.blob_to_kzg_commitment(summed, settings13) == want


# ------------------------------------------------------------------ variable-base MSM (g1_lincomb, BASELINE config 5)
def _synth_scalar(seed, t):
    M = (1 << 64) - 1
    st = 0xB2004844 ^ ((seed * 4096 + t) & M)
    w = b""
    for _ in range(4):
        st = (st + 0x9E3779B97F4A7C15) & M
        z = st
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        w += (z ^ (z >> 31)).to_bytes(8, "big")
    return int.from_bytes(bytes([w[0] & 0x3F]) + w[1:], "big")


@pytest.mark.parametrize("n", [1, 2, 63, 4096, 1 << 14, 1 << 16])
def test_var_msm_synthetic_sizes(lw, settings8, py_setup, n):
    """Points = pseudo-random entries d * 2^(c j) * P_i of the fixed-base table (tau known => known discrete logs)."""
    ms, got = lw.bench_var_msm(n, settings8, iters=1, seed=5)
    dlog = lw.synth_point_dlogs(settings8, n, py_setup.tau)
    acc = 0
    for t in range(n):
        acc = (acc + _synth_scalar(5, t) * dlog[t]) % R
    assert got == bls.g1_compress(bls.g1_mul(bls.G1, acc)), n


def test_g1_lincomb_edge_cases(lw, py_setup):
    rnd = random.Random(77)
    P0, P1 = py_setup.g1[5], py_setup.g1[9]
    enc = lambda p: bytes(96) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")  # noqa: E731
    cases = [
        ([P0, P0, P0], [1, 1, 1]),                       # same point, same bucket -> doubling inside a bucket
        ([P0, bls.g1_neg(P0)], [7, 7]),                  # cancellation
        ([P0, P1, None, P0], [R - 1, R, R + 5, 0]),      # scalars >= r are reduced; infinity point; zero scalar
        ([P0] * 300, [rnd.randrange(1 << 256) for _ in range(300)]),
        ([py_setup.g1[rnd.randrange(4096)] for _ in range(700)], [3] * 700),  # every point in one bucket per window
    ]
    for pts, sc in cases:
        want = bls.g1_compress(bls.g1_sum(bls.g1_mul(p, s % R) for p, s in zip(pts, sc)))
        got = lw.g1_lincomb(b"".join(enc(p) for p in pts), b"".join(s.to_bytes(32, "big") for s in sc), len(pts))
        assert got == want
    with pytest.raises(lw.KzgError):
        lw.g1_lincomb((1).to_bytes(48, "big") + (1).to_bytes(48, "big"), (1).to_bytes(32, "big"), 1)  # (1,1) is not on the curve


# ------------------------------------------------------------------ round-2 regressions
def test_le_mode_unusable_srs_reports_error_not_a_fault(lw):
    """MODE_CKZG_LE with settings whose SRS cannot be re-hydrated (the 4-point setup, zero-padded): every entry
    point -- verification included -- must return C_KZG_ERROR; it used to launch the barycentric kernel on a null
    roots pointer (illegal address, sticky CUDA error for the whole process)."""
    lw.set_option("mode", 1)
    try:
        s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup_4.txt"))
    finally:
        lw.set_option("mode", 0)
    try:
        blob = bytes(kzg.BYTES_PER_BLOB)
        inf = bytes([0xC0]) + bytes(47)
        for call in (lambda: lw.verify_blob_kzg_proof(blob, inf, inf, s),
                     lambda: lw.verify_blob_kzg_proof_batch([blob, blob], [inf, inf], [inf, inf], s),
                     lambda: lw.verify_kzg_proof(inf, bytes(32), bytes(32), inf, s),
                     lambda: lw.verify_batch_phase1(blob + blob, inf + inf, inf + inf, 2, s),
                     lambda: lw.blob_to_kzg_commitment(blob, s),
                     lambda: lw.compute_blob_kzg_proof(blob, inf, s)):
            with pytest.raises(lw.KzgError) as e:
                call()
            assert e.value.code == lw.C_KZG_ERROR
    finally:
        s.free()
    # the process is still healthy: a normal call works afterwards
    lw.set_option("window_bits", 8)
    s2 = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    try:
        assert lw.blob_to_kzg_commitment(blob, s2) == inf
    finally:
        s2.free()


def test_lazy_context_built_once_under_concurrent_first_calls(lw, settings8):
    """Two threads make the first call on the same hand-built KZGSettings at the same time: both get the right
    answer from ONE context (the second waits for the first build instead of racing it)."""
    import threading

    from lambdaworks_kzg_b200.api import CKZGSettings

    g1 = ctypes.create_string_buffer(settings8.g1_values_bytes(), 4096 * 144)
    g2 = ctypes.create_string_buffer(settings8.g2_values_bytes(), 65 * 288)
    s = CKZGSettings(None, ctypes.cast(g1, ctypes.c_void_p), ctypes.cast(g2, ctypes.c_void_p))
    lw.set_option("window_bits", 8)
    blobs = [lw.synth_blob_host(40 + t) for t in range(4)]
    want = [lw.blob_to_kzg_commitment(b, settings8) for b in blobs]
    got, errs = [None] * 4, []

    def work(t):
        try:
            got[t] = lw.blob_to_kzg_commitment(blobs[t], s)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs and got == want
    assert lw.verify_blob_kzg_proof(blobs[0], want[0], lw.compute_blob_kzg_proof(blobs[0], want[0], s), s) is True


def test_device_api_failed_items_are_reported_and_zeroed(lw, settings8, ref):
    """Device-pointer blob proofs with one invalid commitment: its status is C_KZG_ERROR and its proof is zeroed,
    with or without a caller-supplied status buffer; every other item equals the host API's answer."""
    import torch

    n = 5
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    blobs = b"".join(lw.synth_blob_host(900 + k) for k in range(n))
    coms, proofs, _ = lw.commit_and_prove_batch(blobs, n, settings8)
    bad = list(coms)
    bad[2] = bytes(48)                                  # not a valid encoding (no compression flag)
    d_blobs = torch.frombuffer(bytearray(blobs), dtype=torch.uint8).to(dev)
    d_coms = torch.frombuffer(bytearray(b"".join(bad)), dtype=torch.uint8).to(dev)
    for with_status in (True, False):
        d_proofs = torch.full((n * 48,), 0xAB, dtype=torch.uint8, device=dev)
        d_status = torch.full((n,), 7, dtype=torch.int32, device=dev)
        lw.compute_blob_kzg_proof_batch_device(d_proofs.data_ptr(), d_blobs.data_ptr(), d_coms.data_ptr(), n, settings8, st,
                                               d_status.data_ptr() if with_status else 0)
        torch.cuda.synchronize()
        out = bytes(d_proofs.cpu().numpy().tobytes())
        for k in range(n):
            assert out[48 * k: 48 * k + 48] == (bytes(48) if k == 2 else proofs[k]), (with_status, k)
        if with_status:
            assert d_status.cpu().tolist() == [0, 0, lw.C_KZG_ERROR, 0, 0]
    d_c = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    d_status = torch.full((n,), 7, dtype=torch.int32, device=dev)
    lw.blob_to_kzg_commitment_batch_device(d_c.data_ptr(), d_blobs.data_ptr(), n, settings8, st, d_status.data_ptr())
    torch.cuda.synchronize()
    assert bytes(d_c.cpu().numpy().tobytes()) == b"".join(coms) and d_status.cpu().tolist() == [0] * n
    # device-resident verification: same answers as the host call
    d_p = torch.frombuffer(bytearray(b"".join(proofs)), dtype=torch.uint8).to(dev)
    d_good = torch.frombuffer(bytearray(b"".join(coms)), dtype=torch.uint8).to(dev)
    assert lw.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_good.data_ptr(), d_p.data_ptr(), n, settings8) is True
    assert lw.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_good.data_ptr(), d_p.data_ptr(), 1, settings8) is True
    with pytest.raises(lw.KzgError):
        lw.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_coms.data_ptr(), d_p.data_ptr(), n, settings8)
    swapped = torch.frombuffer(bytearray(b"".join([proofs[1], proofs[0]] + proofs[2:])), dtype=torch.uint8).to(dev)
    assert lw.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_good.data_ptr(), swapped.data_ptr(), n, settings8) is False
