"""CPU tier: the PeerDAS / EIP-7594 cell oracle (oracle/py/cells.py) against itself and against the committed known
answers.  The reference implements none of this path and holds no vectors for it (parity unpinned, see the oracle's
header); what pins the restatement is mathematical: cells 0..63 of an extended blob are the blob, every cell is p on
a coset (Horner), a proof computed from the toxic waste equals the explicit quotient MSM, proofs verify under the
universal equation (also with the real pairing), altered inputs do not, and any 64 cells recover the rest."""
import hashlib
import json
import os
import random

import pytest

from oracle.py import bls, cells, kzg
from tests.golden.make_cell_fixtures import PROOF_CELLS, make_blob

R = bls.R
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kats():
    return json.load(open(os.path.join(GOLDEN, "cell_kats.json")))


@pytest.fixture(scope="module")
def material(py_setup):
    o = cells.CellOracle(py_setup, 2)
    blob = make_blob(11, 2)
    cs, ps = o.compute_cells_and_kzg_proofs(blob, cell_subset=PROOF_CELLS)
    c = kzg.DenebMode(py_setup).blob_to_kzg_commitment(blob)
    return o, blob, cs, ps, c


def test_kats_reproduce(py_setup, kats):
    for e in kats:
        o = cells.CellOracle(py_setup, e["mode"])
        blob = make_blob(e["seed"], e["mode"])
        cs, ps = o.compute_cells_and_kzg_proofs(blob, cell_subset=e["proof_cells"])
        assert hashlib.sha256(b"".join(cs)).hexdigest() == e["cells_sha256"]
        assert [p.hex() for p in ps] == e["proofs"]


def test_first_half_of_the_extension_is_the_blob(material):
    o, blob, cs, _, _ = material
    assert b"".join(cs[:64]) == blob


def test_cells_are_coset_evaluations(material):
    o, blob, cs, _, _ = material
    coeffs = o.blob_to_coeffs(blob)
    for ci in (0, 5, 64, 127):
        coset = cells.coset_for_cell(ci)
        ev = o.cell_to_evals(cs[ci])
        for j in (0, 1, 63):
            assert ev[j] == cells.horner(coeffs, coset[j])


def test_commitment_matches_deneb_oracle(material, kats):
    _, _, _, _, c = material
    assert c.hex() == kats[0]["commitment"]


def test_generic_quotient_msm_equals_tau_shortcut(py_setup, material):
    o, blob, _, ps, _ = material
    g = cells.CellOracle(py_setup, 2, generic=True)
    assert bls.g1_compress(g.proof_for_cell(o.blob_to_coeffs(blob), 1)) == ps[PROOF_CELLS.index(1)]


def test_verify_batch(py_setup, material):
    o, blob, cs, ps, c = material
    idx = PROOF_CELLS
    sel = [cs[i] for i in idx]
    assert o.verify_cell_kzg_proof_batch([c] * len(idx), idx, sel, ps) is True
    assert o.verify_cell_kzg_proof_batch([], [], [], []) is True
    bad = list(ps)
    bad[2], bad[3] = bad[3], bad[2]
    assert o.verify_cell_kzg_proof_batch([c] * len(idx), idx, sel, bad) is False
    wrong_cell = list(sel)
    wrong_cell[0] = cs[2]
    assert o.verify_cell_kzg_proof_batch([c] * len(idx), idx, wrong_cell, ps) is False
    other = bls.g1_compress(bls.g1_mul(bls.G1, 5))
    assert o.verify_cell_kzg_proof_batch([c, other] + [c] * (len(idx) - 2), idx, sel, ps) is False
    # the real pairing agrees with the toxic-waste form
    g = cells.CellOracle(py_setup, 2, generic=True)
    assert g.verify_cell_kzg_proof_batch([c] * 2, idx[:2], sel[:2], ps[:2]) is True
    assert g.verify_cell_kzg_proof_batch([c] * 2, idx[:2], sel[:2], [ps[1], ps[0]]) is False


def test_verify_rejects_malformed(material):
    o, blob, cs, ps, c = material
    with pytest.raises(kzg.KzgError):
        o.verify_cell_kzg_proof_batch([c], [128], [cs[0]], [ps[0]])
    noncanon = (R).to_bytes(32, "big") + cs[0][32:]
    with pytest.raises(kzg.KzgError):
        o.verify_cell_kzg_proof_batch([c], [0], [noncanon], [ps[0]])
    with pytest.raises(kzg.KzgError):
        o.verify_cell_kzg_proof_batch([c], [0], [cs[0]], [b"\x00" * 48])


def test_recover(material):
    o, blob, cs, ps, _ = material
    rng = random.Random(3)
    keep = sorted(rng.sample(range(128), 64))
    rc, rp = o.recover_cells_and_kzg_proofs(keep, [cs[i] for i in keep], cell_subset=[100])
    assert rc == cs and rp[0] == ps[PROOF_CELLS.index(100)]
    rc, _ = o.recover_cells_and_kzg_proofs(list(range(64, 128)), cs[64:], want_proofs=False)
    assert rc == cs
    with pytest.raises(kzg.KzgError):
        o.recover_cells_and_kzg_proofs(keep[:63], [cs[i] for i in keep[:63]])
    with pytest.raises(kzg.KzgError):
        o.recover_cells_and_kzg_proofs(keep[::-1], [cs[i] for i in keep[::-1]])


def test_reference_and_le_modes(py_setup):
    for mode in (0, 1):
        o = cells.CellOracle(py_setup, mode)
        blob = make_blob(20 + mode, mode)
        cs, ps = o.compute_cells_and_kzg_proofs(blob, cell_subset=[3, 90])
        coeffs = o.blob_to_coeffs(blob)
        c = bls.g1_compress(bls.g1_mul(bls.G1, cells.horner(coeffs, py_setup.tau)))
        assert o.verify_cell_kzg_proof_batch([c, c], [3, 90], [cs[3], cs[90]], ps) is True
        assert o.verify_cell_kzg_proof_batch([c, c], [90, 3], [cs[3], cs[90]], ps) is False
        assert (b"".join(cs[:64]) == blob) == (mode == 1)   # reference mode: the blob is the coefficient form
