#!/usr/bin/env python3
"""Build the committed golden fixtures from the reference tree.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_fixtures.py

Outputs (all under tests/golden/):
  trusted_setup.txt, trusted_setup_4.txt
        verbatim data fixtures of /root/reference/tests/ (consensus-specs
        *testing* setup, tau = 1337, monomial form) -- inputs, not source.
  ckzg_le_vectors.json + ckzg_le_blobs.bin
        the 208 c-kzg-4844 little-endian-era YAML vectors of
        /root/reference/tests/*/small/*/data.yaml, with the (few, repeated)
        128 KiB blobs de-duplicated into one binary file.  The reference itself
        never loads them (SURVEY.md finding 4); they pin the oracle's Fr/G1/
        SHA/codec arithmetic in LeMode.
  fuzz_corpus.json + fuzz_corpus.bin
        the exact-size libFuzzer seed inputs of /root/reference/fuzz/*/corpus
        (inputs only; there are no expected outputs).
  ref_mode_kats.json
        reference-mode known answers produced by oracle/py (RefMode, both the
        tau shortcut and the generic MSM path must agree) for the valid YAML
        blobs and the re-expressed tests/lib_test.rs cases.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle.py import bls, kzg  # noqa: E402

SUITES = [
    "blob_to_kzg_commitment",
    "compute_kzg_proof",
    "compute_blob_kzg_proof",
    "verify_kzg_proof",
    "verify_blob_kzg_proof",
    "verify_blob_kzg_proof_batch",
]


def unhex(s):
    return bytes.fromhex(s[2:] if s.startswith("0x") else s)


class BlobStore:
    def __init__(self):
        self.index = {}
        self.chunks = []
        self.off = 0

    def put(self, b: bytes):
        k = hashlib.sha1(b).hexdigest()
        if k not in self.index:
            self.index[k] = [self.off, len(b)]
            self.chunks.append(b)
            self.off += len(b)
        return k

    def dump(self, path):
        with open(path, "wb") as f:
            for c in self.chunks:
                f.write(c)


def conv(v, store):
    """hex strings longer than 200 bytes go to the blob store."""
    if isinstance(v, str):
        b = unhex(v)
        if len(b) > 200:
            return {"blob": store.put(b)}
        return b.hex()
    if isinstance(v, list):
        return [conv(x, store) for x in v]
    if isinstance(v, dict):
        return {k: conv(x, store) for k, x in v.items()}
    return v


def main():
    for name in ("trusted_setup.txt", "trusted_setup_4.txt"):
        shutil.copyfile(os.path.join(REF, "tests", name), os.path.join(HERE, name))

    store = BlobStore()
    cases = []
    for suite in SUITES:
        for f in sorted(glob.glob(os.path.join(REF, "tests", suite, "small", "*", "data.yaml"))):
            d = yaml.safe_load(open(f))
            cases.append(
                {
                    "suite": suite,
                    "name": os.path.basename(os.path.dirname(f)),
                    "input": conv(d["input"], store),
                    "output": conv(d["output"], store) if d["output"] is not None else None,
                }
            )
    store.dump(os.path.join(HERE, "ckzg_le_blobs.bin"))
    json.dump({"blobs": store.index, "cases": cases}, open(os.path.join(HERE, "ckzg_le_vectors.json"), "w"), indent=0)
    print("yaml cases:", len(cases), "unique blobs:", len(store.index), "bytes:", store.off)

    # ---- fuzz corpora: exact-size inputs only
    sizes = {
        "blob_to_kzg_commitment": 131072,
        "compute_kzg_proof": 131104,
        "compute_blob_kzg_proof": 131120,
        "verify_kzg_proof": 160,
        "verify_blob_kzg_proof": 131168,
    }
    fstore = BlobStore()
    fz = []
    for suite, ln in sizes.items():
        for f in sorted(glob.glob(os.path.join(REF, "fuzz", suite, "corpus", "*"))):
            if os.path.getsize(f) == ln:
                fz.append({"suite": suite, "name": os.path.basename(f), "data": fstore.put(open(f, "rb").read())})
    # batch: n blobs + n commitments + n proofs, n >= 1
    for f in sorted(glob.glob(os.path.join(REF, "fuzz", "verify_blob_kzg_proof_batch", "corpus", "*"))):
        sz = os.path.getsize(f)
        if sz and sz % 131168 == 0:
            fz.append({"suite": "verify_blob_kzg_proof_batch", "name": os.path.basename(f), "n": sz // 131168, "data": fstore.put(open(f, "rb").read())})
    fstore.dump(os.path.join(HERE, "fuzz_corpus.bin"))
    json.dump({"blobs": fstore.index, "cases": fz}, open(os.path.join(HERE, "fuzz_corpus.json"), "w"), indent=0)
    print("fuzz inputs:", len(fz), "bytes:", fstore.off)

    # ---- reference-mode KATs from the Python oracle
    setup = kzg.parse_setup_text(open(os.path.join(HERE, "trusted_setup.txt")).read())
    fast = kzg.RefMode(setup)
    slow = kzg.RefMode(setup, generic=True)
    blobs_bin = open(os.path.join(HERE, "ckzg_le_blobs.bin"), "rb").read()
    kats = []
    seen = set()
    for c in cases:
        if c["suite"] != "blob_to_kzg_commitment" or "valid_blob" not in c["name"] or "invalid" in c["name"]:
            continue
        key = c["input"]["blob"]["blob"]
        if key in seen:
            continue
        seen.add(key)
        off, ln = store.index[key]
        blob = blobs_bin[off : off + ln]
        com = fast.blob_to_kzg_commitment(blob)
        assert com == slow.blob_to_kzg_commitment(blob), "tau shortcut != generic MSM"
        z2 = (2).to_bytes(32, "big")
        pr, y = fast.compute_kzg_proof(blob, z2)
        bp = fast.compute_blob_kzg_proof(blob, com)
        coeffs = fast.blob_to_coeffs(blob)
        zc = fast.compute_challenge(blob, bls.g1_decompress(com))
        kats.append(
            {
                "name": c["name"].split("_")[-1],
                "blob": key,
                "n_words_ge_r": sum(int.from_bytes(blob[i : i + 32], "big") >= bls.R for i in range(0, len(blob), 32)),
                "commitment": com.hex(),
                "z2_proof": pr.hex(),
                "z2_y": y.hex(),
                "fs_z": zc.to_bytes(32, "big").hex(),
                "fs_y": kzg.horner(coeffs, zc).to_bytes(32, "big").hex(),
                "blob_proof": bp.hex(),
                "verify_blob": fast.verify_blob_kzg_proof(blob, com, bp),
            }
        )
        print("kat", kats[-1]["name"], kats[-1]["commitment"][:16], kats[-1]["n_words_ge_r"])
    json.dump(kats, open(os.path.join(HERE, "ref_mode_kats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
