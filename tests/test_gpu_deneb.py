"""GPU tier: MODE_DENEB -- the final EIP-4844 (mainnet) wire format through the C ABI: big-endian canonical scalars
over the Lagrange SRS, barycentric evaluation, evaluation-form quotient, the final spec's Fiat-Shamir layouts
(the combination the reference left unfinished, /root/reference/src/lib.rs:760-770).

* hash-free outputs: the reference-held YAML vectors with their field elements byte-reversed (tests/_deneb.py);
* hash-dependent outputs: oracle/py DenebMode (restatement of consensus-specs deneb/polynomial-commitments.md)."""
import os
import random

import pytest

from tests._deneb import HASH_FREE_SUITES, rev_fields, to_big_endian
from tests.test_gpu_ckzg_vectors import _lengths_ok, _run, vectors  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B = 4096 * 32


@pytest.fixture(scope="module")
def lw():
    import lambdaworks_kzg_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def dn_settings(lw):
    lw.set_option("mode", 2)
    lw.set_option("window_bits", 8)
    try:
        s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    finally:
        lw.set_option("mode", 0)
    yield s
    s.free()


@pytest.fixture(scope="module")
def dn(py_setup):
    from oracle.py import kzg

    return kzg.DenebMode(py_setup)


def test_hash_free_yaml_vectors_big_endian(lw, dn_settings, vectors):  # noqa: F811
    ran = 0
    for case in vectors:
        suite = case["suite"]
        if suite not in HASH_FREE_SUITES:
            continue
        inp, want = to_big_endian(case)
        if not _lengths_ok(suite, inp):
            assert want is None, case["name"]
            continue
        try:
            got = _run(lw, dn_settings, suite, inp)
        except lw.KzgError as e:
            assert e.code == lw.C_KZG_BADARGS, case["name"]
            got = None
        assert got == want, case["name"]
        ran += 1
    assert ran >= 120


def _blobs(n, seed):
    from oracle.py.bls import R

    rnd = random.Random(seed)
    out = [b"".join(rnd.randrange(R).to_bytes(32, "big") for _ in range(4096)) for _ in range(n - 3)]
    out.append(bytes(B))                                                     # zero polynomial
    out.append((R - 1).to_bytes(32, "big") * 4096)                           # constant r - 1
    out.append(bytes(31) + b"\x01" + bytes(B - 32))                          # a single non-zero evaluation
    return out


def test_blob_proofs_and_verification_vs_oracle(lw, dn_settings, dn):
    blobs = _blobs(6, 1)
    coms = [lw.blob_to_kzg_commitment(b, dn_settings) for b in blobs]
    assert coms == [dn.blob_to_kzg_commitment(b) for b in blobs]
    proofs = [lw.compute_blob_kzg_proof(b, c, dn_settings) for b, c in zip(blobs, coms)]
    assert proofs == [dn.compute_blob_kzg_proof(b, c) for b, c in zip(blobs, coms)]
    for b, c, p in zip(blobs, coms, proofs):
        assert lw.verify_blob_kzg_proof(b, c, p, dn_settings) is True
    assert lw.verify_blob_kzg_proof(blobs[0], coms[0], proofs[1], dn_settings) is False
    # z inside the evaluation domain (compute_quotient_eval_within_domain)
    from oracle.py import kzg

    z = kzg.brp_domain()[5].to_bytes(32, "big")
    assert list(lw.compute_kzg_proof(blobs[0], z, dn_settings)) == list(dn.compute_kzg_proof(blobs[0], z))
    # batch: the whole batch, a swapped pair, the batch challenge itself, the empty batch
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, proofs, dn_settings) is True
    zs = [dn.compute_challenge(b, c) for b, c in zip(blobs, coms)]
    ys = [dn.eval_at(dn.blob_to_evals(b), z) for b, z in zip(blobs, zs)]
    assert lw.debug_batch_challenge(dn_settings) == dn.batch_challenge(coms, zs, ys, proofs)
    swapped = [proofs[1], proofs[0]] + proofs[2:]
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, swapped, dn_settings) is False
    assert dn.verify_blob_kzg_proof_batch(blobs, coms, swapped) is False
    assert lw.verify_blob_kzg_proof_batch([], [], [], dn_settings) is True


def test_invalid_inputs_are_badargs(lw, dn_settings):
    from oracle.py.bls import R

    good = _blobs(4, 2)[0]
    bad = R.to_bytes(32, "big") + good[32:]                  # first word == r
    with pytest.raises(lw.KzgError) as e:
        lw.blob_to_kzg_commitment(bad, dn_settings)
    assert e.value.code == lw.C_KZG_BADARGS
    c = lw.blob_to_kzg_commitment(good, dn_settings)
    with pytest.raises(lw.KzgError) as e:
        lw.compute_kzg_proof(good, b"\xff" * 32, dn_settings)
    assert e.value.code == lw.C_KZG_BADARGS
    with pytest.raises(lw.KzgError) as e:
        lw.compute_blob_kzg_proof(good, b"\x00" * 48, dn_settings)   # not a compressed point
    assert e.value.code == lw.C_KZG_BADARGS
    p = lw.compute_blob_kzg_proof(good, c, dn_settings)
    with pytest.raises(lw.KzgError) as e:
        lw.verify_blob_kzg_proof_batch([good, bad], [c, c], [p, p], dn_settings)
    assert e.value.code == lw.C_KZG_BADARGS
    # the little-endian reading of the same bytes is a different polynomial (words canonical in both byte orders)
    both = b"".join(b"\x00" + good[32 * i + 1: 32 * i + 31] + b"\x00" for i in range(4096))
    assert lw.blob_to_kzg_commitment(rev_fields(both), dn_settings) != lw.blob_to_kzg_commitment(both, dn_settings)


def test_batch_api_large_batch_both_msm_kernels(lw, dn_settings, dn):
    """commit + proof through the batch entry point with the batched-affine kernel forced on, against the XYZZ
    kernel (bit-equal) and against the oracle on a sample."""
    n = 40
    blobs = b"".join(lw.synth_blob_host(k) for k in range(n))    # synthetic words are < 2^254 < r: canonical big-endian
    lw.set_option("msm_ba_min_blobs", 1)
    try:
        c1, p1, st1 = lw.commit_and_prove_batch(blobs, n, dn_settings)
    finally:
        lw.set_option("msm_ba_min_blobs", 5)
    lw.set_option("msm_algo", 0)
    try:
        c0, p0, st0 = lw.commit_and_prove_batch(blobs, n, dn_settings)
    finally:
        lw.set_option("msm_algo", 1)
    assert st0 == [0] * n and st1 == [0] * n and c0 == c1 and p0 == p1
    for k in (0, 17, 39):
        blob = blobs[k * B: (k + 1) * B]
        assert c1[k] == dn.blob_to_kzg_commitment(blob)
        assert p1[k] == dn.compute_blob_kzg_proof(blob, c1[k])
    bl = [blobs[k * B: (k + 1) * B] for k in range(n)]
    assert lw.verify_blob_kzg_proof_batch(bl, c1, p1, dn_settings) is True
