"""GPU tier: the 208 c-kzg-4844 (little-endian era) YAML vectors shipped under
/root/reference/tests/*/small, through the C ABI in MODE_CKZG_LE (SURVEY §8f.1,
App. B).  Inputs whose lengths a C ABI cannot express (too few / too many bytes)
must be expected errors and are rejected here, as a binding layer would."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B = 4096 * 32


@pytest.fixture(scope="module")
def lw():
    import lambdaworks_kzg_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def le_settings(lw):
    lw.set_option("mode", 1)
    lw.set_option("window_bits", 8)
    try:
        s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    finally:
        lw.set_option("mode", 0)
    yield s
    s.free()


@pytest.fixture(scope="module")
def vectors():
    meta = json.load(open(os.path.join(GOLDEN, "ckzg_le_vectors.json")))
    raw = open(os.path.join(GOLDEN, "ckzg_le_blobs.bin"), "rb").read()

    def resolve(v):
        if isinstance(v, dict) and len(v) == 1 and isinstance(v.get("blob"), str):
            off, ln = meta["blobs"][v["blob"]]
            return raw[off: off + ln]
        if isinstance(v, dict):
            return {k: resolve(x) for k, x in v.items()}
        if isinstance(v, list):
            return [resolve(x) for x in v]
        if isinstance(v, str):
            return bytes.fromhex(v)
        return v

    return [dict(c, input=resolve(c["input"]), output=resolve(c["output"])) for c in meta["cases"]]


def _lengths_ok(suite, inp):
    chk = {"blob": B, "commitment": 48, "proof": 48, "z": 32, "y": 32}
    for k, v in inp.items():
        if k in chk and len(v) != chk[k]:
            return False
        if k == "blobs" and any(len(x) != B for x in v):
            return False
        if k in ("commitments", "proofs") and any(len(x) != 48 for x in v):
            return False
    if suite == "verify_blob_kzg_proof_batch" and not (len(inp["blobs"]) == len(inp["commitments"]) == len(inp["proofs"])):
        return False
    return True


def _run(lw, s, suite, inp):
    if suite == "blob_to_kzg_commitment":
        return lw.blob_to_kzg_commitment(inp["blob"], s)
    if suite == "compute_kzg_proof":
        p, y = lw.compute_kzg_proof(inp["blob"], inp["z"], s)
        return [p, y]
    if suite == "compute_blob_kzg_proof":
        return lw.compute_blob_kzg_proof(inp["blob"], inp["commitment"], s)
    if suite == "verify_kzg_proof":
        return lw.verify_kzg_proof(inp["commitment"], inp["z"], inp["y"], inp["proof"], s)
    if suite == "verify_blob_kzg_proof":
        return lw.verify_blob_kzg_proof(inp["blob"], inp["commitment"], inp["proof"], s)
    if suite == "verify_blob_kzg_proof_batch":
        return lw.verify_blob_kzg_proof_batch(inp["blobs"], inp["commitments"], inp["proofs"], s)
    raise AssertionError(suite)


def test_all_yaml_vectors_through_the_gpu(lw, le_settings, vectors):
    assert len(vectors) == 208
    ran = 0
    for case in vectors:
        suite, inp, want = case["suite"], case["input"], case["output"]
        if not _lengths_ok(suite, inp):
            assert want is None, case["name"]  # length errors: rejected before the C ABI
            continue
        try:
            got = _run(lw, le_settings, suite, inp)
        except lw.KzgError as e:
            assert e.code == lw.C_KZG_BADARGS, case["name"]  # c-kzg reports invalid input as BADARGS
            got = None
        assert got == want, case["name"]
        ran += 1
    assert ran == 177  # the other 31 cases are wrong-length inputs (not expressible at a C ABI)


def test_lagrange_setup_layout(lw, le_settings, py_setup):
    """g1_values holds the Lagrange-form SRS in bit-reversed order, as c-kzg's loader stores it:
    L_i = [l_i(tau)]G with l_i the Lagrange polynomial of the domain point w^brp(i)."""
    from oracle.py import bls, kzg

    dom = kzg.brp_domain()
    tau, R = py_setup.tau, bls.R
    g1 = le_settings.g1_values_bytes()
    zn = (pow(tau, 4096, R) - 1) * pow(4096, -1, R) % R
    for i in (0, 1, 2, 4095):
        li = zn * dom[i] % R * pow((tau - dom[i]) % R, -1, R) % R
        x, y = bls.g1_mul(bls.G1, li)
        limbs = lambda v: b"".join(((v >> (64 * (5 - k))) & (2**64 - 1)).to_bytes(8, "little") for k in range(6))  # noqa: E731
        assert g1[144 * i: 144 * i + 96] == limbs(x) + limbs(y), i
