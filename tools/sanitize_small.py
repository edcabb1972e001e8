#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_kzg_b200 as lw

setup = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt")
for mode in (0, 1):
    lw.set_option("mode", mode)
    lw.set_option("window_bits", 4)
    s = lw.load_trusted_setup_file(setup)
    n = 3
    blobs = [lw.synth_blob_host(k) for k in range(n)]
    if mode == 1:  # little-endian canonical: move the cleared byte to the top of each word
        blobs = [b"".join(b[i:i + 32][::-1] for i in range(0, len(b), 32)) for b in blobs]
    coms, proofs, st = lw.commit_and_prove_batch(b"".join(blobs), n, s)
    assert st == [0] * n
    # the same batch through the batched-affine MSM kernel (normally used from 256 blobs up), plus degenerate blobs
    lw.set_option("msm_ba_min_blobs", 1)
    for variant in (0, 3):
        lw.set_option("msm_ba_variant", variant)
        coms2, proofs2, st2 = lw.commit_and_prove_batch(b"".join(blobs) + bytes(131072) + (bytes(31) + b"\x01") * 4096, n + 2, s)
        assert st2 == [0] * (n + 2) and coms2[:n] == coms and proofs2[:n] == proofs
    lw.set_option("msm_ba_variant", 0)
    lw.set_option("msm_ba_min_blobs", 5)
    assert lw.verify_blob_kzg_proof_batch(blobs, coms, proofs, s) is True
    assert lw.verify_blob_kzg_proof(blobs[0], coms[0], proofs[0], s) is True
    z = bytes(31) + b"\x07" if mode == 0 else b"\x07" + bytes(31)
    p, y = lw.compute_kzg_proof(blobs[1], z, s)
    assert lw.verify_kzg_proof(coms[1], z, y, p, s) is True
    assert lw.verify_kzg_proof(coms[1], z, y, proofs[0], s) is False
    # the large-batch verification pipeline at a small size: group SHA kernel (n > 64, n % 8 != 0), several staging
    # super-batches with a partial last one, odd chunks, chunk-by-chunk batch challenge; blob proofs for given commitments
    n2 = int(os.environ.get("SAN_VERIFY_BLOBS", "70"))
    blobs2 = [lw.synth_blob_host(100 + k) for k in range(n2)]
    if mode == 1:
        blobs2 = [b"".join(b[i:i + 32][::-1] for i in range(0, len(b), 32)) for b in blobs2]
    c2, p2, st2 = lw.commit_and_prove_batch(b"".join(blobs2), n2, s)
    assert st2 == [0] * n2
    p3, st3 = lw.compute_blob_kzg_proof_batch(b"".join(blobs2), b"".join(c2), n2, s)
    assert st3 == [0] * n2 and p3 == p2
    lw.set_option("verify_super_blobs", 32)
    lw.set_option("chunk_blobs", 20)
    assert lw.verify_blob_kzg_proof_batch(blobs2, c2, p2, s) is True
    bad = list(p2); bad[-1] = c2[0]
    assert lw.verify_blob_kzg_proof_batch(blobs2, c2, bad, s) is False
    lw.set_option("verify_super_blobs", 16384)
    lw.set_option("chunk_blobs", 256)
    assert lw.verify_blob_kzg_proof_batch(blobs2, c2, p2, s) is True
    s.free()
lw.set_option("mode", 0)
pts = bytes(96) + b"".join(bytes.fromhex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1") for _ in range(40))
sc = b"".join((k * 2654435761 % (1 << 200)).to_bytes(32, "big") for k in range(41))
print("lincomb", lw.g1_lincomb(pts, sc, 41).hex()[:16], "launches", lw.kernel_launches())
