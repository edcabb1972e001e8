"""A handful of small calls through the new kernels, for compute-sanitizer --tool racecheck / synccheck (tiny tables)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_kzg_b200 as lw

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
lw.set_option("mode", 2)
lw.set_option("window_bits", 4)
lw.set_option("cell_window_bits", 4)
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
rng = random.Random(3)
blob = b"".join(rng.randrange(R).to_bytes(32, "big") for _ in range(4096))
cells, proofs = lw.compute_cells_and_kzg_proofs(blob, s)
if os.environ.get("ONLY_COMPUTE"):
    keep = list(range(0, 128, 2))
    rc, rp = lw.recover_cells_and_kzg_proofs(keep, [cells[i] for i in keep], s)
    assert rc == cells and rp == proofs
    print("SANITIZE_CELLS_OK (compute + recover only)")
    sys.exit(0)
com = lw.blob_to_kzg_commitment(blob, s)
idx = [0, 5, 64, 127]
assert lw.verify_cell_kzg_proof_batch([com] * 4, idx, [cells[i] for i in idx], [proofs[i] for i in idx], s) is True
keep = list(range(0, 128, 2))
rc, rp = lw.recover_cells_and_kzg_proofs(keep, [cells[i] for i in keep], s)
assert rc == cells and rp == proofs
p, y = lw.compute_kzg_proof(blob, bytes(31) + b"\x07", s)
assert lw.verify_kzg_proof(com, bytes(31) + b"\x07", y, p, s) is True
pr = lw.compute_blob_kzg_proof(blob, com, s)
assert lw.verify_blob_kzg_proof(blob, com, pr, s) is True
assert lw.verify_blob_kzg_proof_batch([blob] * 3, [com] * 3, [pr] * 3, s) is True
print("SANITIZE_CELLS_OK")
