#!/usr/bin/env python3
"""Writes tests/golden/cell_kats.json: known answers for the PeerDAS / EIP-7594 cell calls, produced by the Python
oracle (oracle/py/cells.py) WITHOUT FK20 -- every proof is [(p(tau) - I_k(tau)) / (tau^64 - h_k^64)]G from the toxic
waste of tests/golden/trusted_setup.txt, one of them cross-checked against the explicit quotient MSM.  The reference
holds no cell vectors (it does not implement this path), so these pin the restatement, not the reference.

Blob generator: random.Random(seed), 4096 x randrange(r), each written in the mode's byte order."""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.py import bls, cells, kzg  # noqa: E402

PROOF_CELLS = [0, 1, 63, 64, 100, 127]


def make_blob(seed: int, mode: int) -> bytes:
    rng = random.Random(seed)
    return b"".join(rng.randrange(bls.R).to_bytes(32, "little" if mode == 1 else "big") for _ in range(4096))


def main():
    setup = kzg.parse_setup_text(open(os.path.join(HERE, "trusted_setup.txt")).read())
    out = []
    for mode, seed in ((2, 11), (2, 12), (1, 13), (0, 14)):
        o = cells.CellOracle(setup, mode)
        blob = make_blob(seed, mode)
        cs, ps = o.compute_cells_and_kzg_proofs(blob, cell_subset=PROOF_CELLS)
        coeffs = o.blob_to_coeffs(blob)
        commitment = bls.g1_compress(bls.g1_mul(bls.G1, cells.horner(coeffs, setup.tau)))
        entry = {"mode": mode, "seed": seed, "cells_sha256": hashlib.sha256(b"".join(cs)).hexdigest(), "commitment": commitment.hex(),
                 "proof_cells": PROOF_CELLS, "proofs": [p.hex() for p in ps]}
        if mode == 2 and seed == 11:
            g = cells.CellOracle(setup, mode, generic=True)
            assert bls.g1_compress(g.proof_for_cell(coeffs, 100)) == ps[PROOF_CELLS.index(100)]
            entry["generic_msm_checked_cell"] = 100
        out.append(entry)
    json.dump(out, open(os.path.join(HERE, "cell_kats.json"), "w"), indent=1)
    print("wrote cell_kats.json:", len(out), "entries")


if __name__ == "__main__":
    main()
