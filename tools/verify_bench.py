#!/usr/bin/env python3
"""Timing of the verification entry points (BASELINE config 3) and single-call latencies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

n = int(os.environ.get("NB", "4096"))
lw.set_option("window_bits", int(os.environ.get("WB", "13")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
blobs_d = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(blobs_d.data_ptr(), 0, n, 0)
torch.cuda.synchronize()
blobs = bytes(blobs_d.cpu().numpy().tobytes())
t = time.perf_counter(); coms, proofs, st = lw.commit_and_prove_batch(blobs, n, s); dt = time.perf_counter() - t
print("commit+prove host API n=%d: %.1f ms (%.0f blobs/s)" % (n, dt * 1e3, n / dt), flush=True)
B = 131072
bl = [blobs[i * B:(i + 1) * B] for i in range(n)]
for rep in range(3):
    t = time.perf_counter(); ok = lw.verify_blob_kzg_proof_batch(bl, coms, proofs, s); dt = time.perf_counter() - t
    print("verify_blob_kzg_proof_batch n=%d -> %s: %.1f ms (%.0f blobs/s)" % (n, ok, dt * 1e3, n / dt), flush=True)
# the same call from pinned host memory through raw pointers (no Python joins, asynchronous H2D)
hb = torch.frombuffer(bytearray(blobs), dtype=torch.uint8).pin_memory()
hc = torch.frombuffer(bytearray(b"".join(coms)), dtype=torch.uint8).pin_memory()
hp = torch.frombuffer(bytearray(b"".join(proofs)), dtype=torch.uint8).pin_memory()
for rep in range(3):
    t = time.perf_counter(); ok = lw.verify_blob_kzg_proof_batch_ptr(hb.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, s); dt = time.perf_counter() - t
    print("verify_blob_kzg_proof_batch (pinned, pointer API) n=%d -> %s: %.1f ms (%.0f blobs/s)" % (n, ok, dt * 1e3, n / dt), flush=True)
bad = list(proofs); bad[n // 2 - 1] = coms[0]
t = time.perf_counter(); ok = lw.verify_blob_kzg_proof_batch(bl, coms, bad, s); dt = time.perf_counter() - t
print("verify batch with proof #%d replaced -> %s: %.1f ms" % (n // 2 - 1, ok, dt * 1e3), flush=True)
z5 = bytes(31) + b"\x05"
p5, y5 = lw.compute_kzg_proof(bl[0], z5, s)
assert lw.verify_kzg_proof(coms[0], z5, y5, p5, s) is True
for name, fn in [("blob_to_kzg_commitment", lambda: lw.blob_to_kzg_commitment(bl[0], s)),
                 ("compute_blob_kzg_proof", lambda: lw.compute_blob_kzg_proof(bl[0], coms[0], s)),
                 ("compute_kzg_proof", lambda: lw.compute_kzg_proof(bl[0], bytes(31) + b"\x05", s)),
                 ("verify_blob_kzg_proof", lambda: lw.verify_blob_kzg_proof(bl[0], coms[0], proofs[0], s)),
                 ("verify_kzg_proof", lambda: lw.verify_kzg_proof(coms[0], z5, y5, p5, s))]:
    fn()
    t = time.perf_counter()
    for _ in range(5):
        r = fn()
    print("%-26s %.3f ms/call" % (name, (time.perf_counter() - t) / 5 * 1e3), flush=True)
