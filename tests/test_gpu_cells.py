"""GPU tier: PeerDAS / EIP-7594 cells and FK20 cell proofs through the C ABI (include/lwkzg.h part 3)
against oracle/py/cells.py, which computes every proof WITHOUT FK20 (toxic-waste quotient, one cross-checked by the
explicit MSM) -- the reference has no counterpart (/root/reference/src/srs.rs:274 reads 2 of the 65 G2 points)."""
import hashlib
import json
import os
import random

import pytest

from tests.golden.make_cell_fixtures import make_blob

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B = 4096 * 32
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@pytest.fixture(scope="module")
def lw():
    import lambdaworks_kzg_b200 as m

    m.load_library()
    return m


def _load(lw, mode):
    lw.set_option("mode", mode)
    lw.set_option("window_bits", 8)
    lw.set_option("cell_window_bits", 8)
    try:
        return lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    finally:
        lw.set_option("mode", 0)
        lw.set_option("window_bits", 13)
        lw.set_option("cell_window_bits", 13)


@pytest.fixture(scope="module")
def s2(lw):
    s = _load(lw, 2)
    yield s
    s.free()


@pytest.fixture(scope="module")
def o2(py_setup):
    from oracle.py import cells

    return cells.CellOracle(py_setup, 2)


@pytest.fixture(scope="module")
def blob11(lw, s2):
    blob = make_blob(11, 2)
    cs, ps = lw.compute_cells_and_kzg_proofs(blob, s2)
    return blob, cs, ps


def test_kats(lw, s2, blob11):
    kats = json.load(open(os.path.join(GOLDEN, "cell_kats.json")))
    e = kats[0]
    blob, cs, ps = blob11
    assert hashlib.sha256(b"".join(cs)).hexdigest() == e["cells_sha256"]
    assert [ps[i].hex() for i in e["proof_cells"]] == e["proofs"]
    assert lw.blob_to_kzg_commitment(blob, s2).hex() == e["commitment"]
    e = kats[1]
    cs, ps = lw.compute_cells_and_kzg_proofs(make_blob(e["seed"], 2), s2)
    assert hashlib.sha256(b"".join(cs)).hexdigest() == e["cells_sha256"]
    assert [ps[i].hex() for i in e["proof_cells"]] == e["proofs"]


def test_all_cells_and_proofs_vs_oracle(lw, s2, o2, blob11):
    from oracle.py import bls

    blob, cs, ps = blob11
    assert b"".join(cs[:64]) == blob
    coeffs = o2.blob_to_coeffs(blob)
    want_cells = [o2.cell_to_bytes(c) for c in o2.cells_from_coeffs(coeffs)]
    assert cs == want_cells
    for k in range(128):
        assert ps[k] == bls.g1_compress(o2.proof_for_cell(coeffs, k)), "proof %d" % k
    # cells only / proofs only give the same bytes
    c_only, none = lw.compute_cells_and_kzg_proofs(blob, s2, want_proofs=False)
    assert c_only == cs and none == []
    none, p_only = lw.compute_cells_and_kzg_proofs(blob, s2, want_cells=False)
    assert p_only == ps and none == []


def test_edge_blobs(lw, s2, o2):
    from oracle.py import bls

    zero = bytes(B)
    cs, ps = lw.compute_cells_and_kzg_proofs(zero, s2)
    assert all(c == bytes(2048) for c in cs)
    inf = bytes([0xC0]) + bytes(47)
    assert all(p == inf for p in ps)
    # constant polynomial: every evaluation equal, every quotient zero
    const = (5).to_bytes(32, "big") * 4096
    cs, ps = lw.compute_cells_and_kzg_proofs(const, s2)
    assert all(c == (5).to_bytes(32, "big") * 64 for c in cs) and all(p == inf for p in ps)
    # largest canonical values, and a blob with a single non-zero evaluation
    for blob in ((R - 1).to_bytes(32, "big") * 4096, bytes(32 * 4095) + (1).to_bytes(32, "big"), (1).to_bytes(32, "big") + bytes(32 * 4095)):
        cs, ps = lw.compute_cells_and_kzg_proofs(blob, s2)
        coeffs = o2.blob_to_coeffs(blob)
        assert cs == [o2.cell_to_bytes(c) for c in o2.cells_from_coeffs(coeffs)]
        for k in (0, 37, 64, 127):
            assert ps[k] == bls.g1_compress(o2.proof_for_cell(coeffs, k))
    # a word >= r is rejected and nothing is written
    bad = bytes(32 * 7) + R.to_bytes(32, "big") + bytes(32 * 4088)
    with pytest.raises(lw.KzgError) as ei:
        lw.compute_cells_and_kzg_proofs(bad, s2)
    assert ei.value.code == lw.C_KZG_BADARGS


def test_batch_and_device_api_equal_single_calls(lw, s2, blob11):
    import torch

    n = 5
    blobs = [make_blob(30 + i, 2) for i in range(n)]
    blobs[3] = bytes(32 * 9) + (R + 1).to_bytes(32, "big") + bytes(32 * 4086)   # invalid item in the middle
    singles = [lw.compute_cells_and_kzg_proofs(b, s2) if i != 3 else None for i, b in enumerate(blobs)]
    lw.set_option("cell_chunk_blobs", 2)   # three passes: 2 + 2 + 1
    try:
        cells, proofs, st = lw.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n, s2)
    finally:
        lw.set_option("cell_chunk_blobs", 888)
    assert st == [0, 0, 0, lw.C_KZG_BADARGS, 0]
    for i in range(n):
        if i == 3:
            continue
        assert cells[i * 262144: (i + 1) * 262144] == b"".join(singles[i][0])
        assert proofs[i * 6144: (i + 1) * 6144] == b"".join(singles[i][1])
    # device pointers
    d_blobs = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
    d_cells = torch.zeros(n * 262144, dtype=torch.uint8, device="cuda")
    d_proofs = torch.zeros(n * 6144, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lw.compute_cells_and_kzg_proofs_batch_device(d_cells.data_ptr(), d_proofs.data_ptr(), d_blobs.data_ptr(), n, s2,
                                                 torch.cuda.current_stream().cuda_stream, d_st.data_ptr())
    torch.cuda.synchronize()
    assert d_st.tolist() == st
    hc, hp = bytes(d_cells.cpu().numpy()), bytes(d_proofs.cpu().numpy())
    for i in range(n):
        if i == 3:
            assert hc[i * 262144: (i + 1) * 262144] == bytes(262144) and hp[i * 6144: (i + 1) * 6144] == bytes(6144)
        else:
            assert hc[i * 262144: (i + 1) * 262144] == cells[i * 262144: (i + 1) * 262144]
            assert hp[i * 6144: (i + 1) * 6144] == proofs[i * 6144: (i + 1) * 6144]


def test_verify_cell_kzg_proof_batch(lw, s2, o2, blob11):
    blob, cs, ps = blob11
    c = lw.blob_to_kzg_commitment(blob, s2)
    blob_b = make_blob(12, 2)
    cs_b, ps_b = lw.compute_cells_and_kzg_proofs(blob_b, s2)
    c_b = lw.blob_to_kzg_commitment(blob_b, s2)
    assert lw.verify_cell_kzg_proof_batch([], [], [], [], s2) is True
    # one cell, a few cells, all cells of one blob
    assert lw.verify_cell_kzg_proof_batch([c], [5], [cs[5]], [ps[5]], s2) is True
    idx = [0, 1, 63, 64, 100, 127]
    assert lw.verify_cell_kzg_proof_batch([c] * 6, idx, [cs[i] for i in idx], [ps[i] for i in idx], s2) is True
    assert o2.verify_cell_kzg_proof_batch([c] * 6, idx, [cs[i] for i in idx], [ps[i] for i in idx]) is True
    assert lw.verify_cell_kzg_proof_batch([c] * 128, list(range(128)), cs, ps, s2) is True
    # two blobs interleaved, repeated cells
    coms = [c, c_b, c, c_b, c_b, c]
    idx2 = [7, 7, 99, 3, 3, 7]
    cells2 = [cs[7], cs_b[7], cs[99], cs_b[3], cs_b[3], cs[7]]
    proofs2 = [ps[7], ps_b[7], ps[99], ps_b[3], ps_b[3], ps[7]]
    assert lw.verify_cell_kzg_proof_batch(coms, idx2, cells2, proofs2, s2) is True
    # negatives: swapped proofs, wrong cell, wrong index, wrong commitment -- the oracle agrees on each
    cases = [
        (coms, idx2, cells2, [proofs2[1], proofs2[0]] + proofs2[2:]),
        (coms, idx2, [cs[8]] + cells2[1:], proofs2),
        (coms, [8] + idx2[1:], cells2, proofs2),
        ([c_b] + coms[1:], idx2, cells2, proofs2),
    ]
    for a, b, d, e in cases:
        assert lw.verify_cell_kzg_proof_batch(a, b, d, e, s2) is False
        assert o2.verify_cell_kzg_proof_batch(a, b, d, e) is False
    # malformed inputs
    for a, b, d, e in [
        ([c], [128], [cs[0]], [ps[0]]),
        ([c], [0], [R.to_bytes(32, "big") + cs[0][32:]], [ps[0]]),
        ([c], [0], [cs[0]], [bytes(48)]),
        ([bytes([0x80]) + bytes(47)], [0], [cs[0]], [ps[0]]),
    ]:
        with pytest.raises(lw.KzgError) as ei:
            lw.verify_cell_kzg_proof_batch(a, b, d, e, s2)
        assert ei.value.code == lw.C_KZG_BADARGS


def test_recover_cells_and_kzg_proofs(lw, s2, blob11):
    blob, cs, ps = blob11
    rng = random.Random(9)
    for keep in (sorted(rng.sample(range(128), 64)), list(range(64, 128)), list(range(0, 128, 2)), sorted(rng.sample(range(128), 101)), list(range(128))):
        rc, rp = lw.recover_cells_and_kzg_proofs(keep, [cs[i] for i in keep], s2)
        assert rc == cs and rp == ps
    keep = list(range(64))
    rc, none = lw.recover_cells_and_kzg_proofs(keep, [cs[i] for i in keep], s2, want_proofs=False)
    assert rc == cs and none == []
    # too few, unsorted, repeated, out of range, inconsistent, non-canonical
    for idx, sel in [
        (list(range(63)), cs[:63]),
        ([1, 0] + list(range(2, 64)), [cs[1], cs[0]] + cs[2:64]),
        ([0, 0] + list(range(2, 64)), [cs[0], cs[0]] + cs[2:64]),
        (list(range(63)) + [128], cs[:64]),
        (list(range(65)), cs[:64] + [cs[3]]),
        (list(range(64)), [R.to_bytes(32, "big") + cs[0][32:]] + cs[1:64]),
    ]:
        with pytest.raises(lw.KzgError) as ei:
            lw.recover_cells_and_kzg_proofs(idx, sel, s2)
        assert ei.value.code == lw.C_KZG_BADARGS, idx[:3]


@pytest.mark.parametrize("mode", [0, 1])
def test_reference_and_le_modes(lw, py_setup, mode):
    from oracle.py import bls, cells

    o = cells.CellOracle(py_setup, mode)
    kat = [e for e in json.load(open(os.path.join(GOLDEN, "cell_kats.json"))) if e["mode"] == mode][0]
    s = _load(lw, mode)
    try:
        blob = make_blob(kat["seed"], mode)
        cs, ps = lw.compute_cells_and_kzg_proofs(blob, s)
        assert hashlib.sha256(b"".join(cs)).hexdigest() == kat["cells_sha256"]
        assert [ps[i].hex() for i in kat["proof_cells"]] == kat["proofs"]
        c = lw.blob_to_kzg_commitment(blob, s)
        assert c.hex() == kat["commitment"]
        idx = [2, 64, 77]
        args = ([c] * 3, idx, [cs[i] for i in idx], [ps[i] for i in idx])
        assert lw.verify_cell_kzg_proof_batch(*args, s) is True and o.verify_cell_kzg_proof_batch(*args) is True
        assert lw.verify_cell_kzg_proof_batch([c] * 3, idx, [cs[i] for i in idx], [ps[64], ps[2], ps[77]], s) is False
        keep = list(range(1, 128, 2))
        rc, rp = lw.recover_cells_and_kzg_proofs(keep, [cs[i] for i in keep], s)
        assert rc == cs and rp == ps
        if mode == 0:   # every failure is C_KZG_ERROR in reference mode; blob words >= r are reduced, not rejected
            with pytest.raises(lw.KzgError) as ei:
                lw.verify_cell_kzg_proof_batch([c], [128], [cs[0]], [ps[0]], s)
            assert ei.value.code == lw.C_KZG_ERROR
            big = R.to_bytes(32, "big") + blob[32:]
            cs2, _ = lw.compute_cells_and_kzg_proofs(big, s, want_proofs=False)
            assert cs2 == [o.cell_to_bytes(x) for x in o.cells_from_coeffs(o.blob_to_coeffs(big))]
    finally:
        s.free()


def test_full_batch_properties(lw, s2):
    """64 blobs at once (8192 cell proofs): every (blob, cell) pair of a sample verifies in ONE batched check over 64
    commitments, and a single swapped proof flips it."""
    n = 64
    blobs = b"".join(lw.synth_blob_host(k) for k in range(n))
    cells, proofs, st = lw.compute_cells_and_kzg_proofs_batch(blobs, n, s2)
    assert st == [0] * n
    coms, _ = lw.blob_to_kzg_commitment_batch(blobs, n, s2)
    rng = random.Random(4)
    pick = [(b, rng.randrange(128)) for b in range(n) for _ in range(4)]
    cell = lambda b, i: cells[(b * 128 + i) * 2048: (b * 128 + i + 1) * 2048]
    proof = lambda b, i: proofs[(b * 128 + i) * 48: (b * 128 + i + 1) * 48]
    args = ([coms[b] for b, _ in pick], [i for _, i in pick], [cell(b, i) for b, i in pick], [proof(b, i) for b, i in pick])
    assert lw.verify_cell_kzg_proof_batch(*args, s2) is True
    pr = list(args[3])
    pr[17], pr[18] = pr[18], pr[17]
    assert lw.verify_cell_kzg_proof_batch(args[0], args[1], args[2], pr, s2) is False


def test_fk20_stages_against_the_exponent_model(lw, s2, o2):
    """Every intermediate of the FK20 pipeline against a model in the exponent (tau known): the 8192 MSM scalars, the
    8192 transformed SRS points, the 128 MSM results and the inverse G1 FFT -- so a failure names its stage."""
    from oracle.py import bls, cells
    from oracle.py.kzg import _bitrev

    Rr, tau = bls.R, 1337
    blob = make_blob(41, 2)
    f = o2.blob_to_coeffs(blob)
    scalars, hhat, h, fk = lw.debug_cell_stages(blob, s2)
    nu = pow(cells.root_of_unity(8192), 64, Rr)
    sp = [pow(tau, k, Rr) for k in range(4096)]
    inv128 = bls.fr_inv(128)
    X, C = [], []
    for b in range(64):
        X.append(cells.fft([sp[64 * (62 - v) + b] if v <= 62 else 0 for v in range(128)], nu))
        c = [0] * 128
        c[0] = f[64 * 63 + b]
        for u in range(66, 128):
            c[u] = f[64 * (u - 65) + b]
        C.append(cells.fft(c, nu))
    assert all(scalars[j * 64 + b] == C[b][j] * inv128 % Rr for j in range(128) for b in range(64))
    rng = random.Random(2)
    for j, b in [(0, 0), (64, 5), (127, 63)] + [(rng.randrange(128), rng.randrange(64)) for _ in range(40)]:
        assert fk[j * 64 + b] == bls.g1_mul(bls.G1, X[b][j]), (j, b)
    hh = [sum(C[b][j] * X[b][j] for b in range(64)) * inv128 % Rr for j in range(128)]
    assert all(hhat[j] == bls.g1_compress(bls.g1_mul(bls.G1, hh[j])) for j in range(128))
    H = cells.fft(hh, bls.fr_inv(nu))
    assert H[63] == 0
    for p in range(128):
        want = bls.g1_mul(bls.G1, H[_bitrev(p, 7)]) if p % 2 == 0 else None
        assert h[p] == bls.g1_compress(want), p


def test_large_batch_kernel_equals_small_batch_kernel(lw, s2):
    """From 224 blobs up the FK20 MSMs run on the batched-affine kernel in its segmented form; below, one warp per
    (blob, frequency).  Same bytes either way, including degenerate blobs inside the large batch."""
    n = 230
    blobs = [lw.synth_blob_host(900 + k) for k in range(n)]
    blobs[5] = bytes(B)                                   # zero polynomial: every proof infinity
    blobs[77] = (9).to_bytes(32, "big") * 4096            # constant
    blobs[229] = bytes(32 * 4095) + (1).to_bytes(32, "big")
    _, big, st = lw.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n, s2, want_cells=False)
    assert st == [0] * n
    lw.set_option("cell_chunk_blobs", 100)               # passes of 100 + 100 + 30 blobs: the warp-per-frequency kernel
    try:
        _, small, st = lw.compute_cells_and_kzg_proofs_batch(b"".join(blobs), n, s2, want_cells=False)
    finally:
        lw.set_option("cell_chunk_blobs", 888)
    assert st == [0] * n and big == small
    inf = bytes([0xC0]) + bytes(47)
    assert big[5 * 6144: 6 * 6144] == inf * 128 and big[77 * 6144: 78 * 6144] == inf * 128


def test_cells_in_library_multi_device_byte_equality():
    """lwkzg_set_devices: a cell batch sharded over every visible GPU from one process equals the single-device bytes."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "multi_device_cells_check.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "MULTI_DEVICE_CELLS_OK" in p.stdout


def test_cells_are_linear_and_mode_independent(lw, s2, o2):
    """Two properties that need neither the toxic waste nor the oracle's proof code: (1) proofs are linear in the blob
    -- pi_i(a + b) = pi_i(a) + pi_i(b) as group elements, cells add field-wise; (2) the same polynomial given in
    evaluation form (MODE_DENEB) and in coefficient form (MODE_REFERENCE) has the same cells and the same proofs."""
    from oracle.py import bls

    a = make_blob(51, 2)
    b = make_blob(52, 2)
    wa = [int.from_bytes(a[i: i + 32], "big") for i in range(0, B, 32)]
    wb = [int.from_bytes(b[i: i + 32], "big") for i in range(0, B, 32)]
    c = b"".join(((x + y) % R).to_bytes(32, "big") for x, y in zip(wa, wb))
    (ca, pa), (cb, pb), (cc, pc) = (lw.compute_cells_and_kzg_proofs(x, s2) for x in (a, b, c))
    for i in range(128):
        ea = [int.from_bytes(ca[i][j: j + 32], "big") for j in range(0, 2048, 32)]
        eb = [int.from_bytes(cb[i][j: j + 32], "big") for j in range(0, 2048, 32)]
        assert cc[i] == b"".join(((x + y) % R).to_bytes(32, "big") for x, y in zip(ea, eb))
        assert pc[i] == bls.g1_compress(bls.g1_add(bls.g1_decompress(pa[i]), bls.g1_decompress(pb[i]))), i
    # coefficient form of a, through the reference-mode settings
    coeffs = o2.blob_to_coeffs(a)
    s0 = _load(lw, 0)
    try:
        c0, p0 = lw.compute_cells_and_kzg_proofs(b"".join(v.to_bytes(32, "big") for v in coeffs), s0)
    finally:
        s0.free()
    assert c0 == ca and p0 == pa
