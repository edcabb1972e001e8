#!/usr/bin/env python3
"""lwkzg_set_devices: ONE process, one C call per batch, every visible GPU.  The sharded calls must
return the bytes of the single-device calls, the sharded batched verification the same boolean.  Prints
MULTI_DEVICE_CHECK_OK on success."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

g = torch.cuda.device_count()
assert g >= 2, "needs at least two GPUs"
n = int(os.environ.get("NB", "1536"))
B = 131072
lw.set_option("window_bits", int(os.environ.get("WB", "13")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
blobs = b"".join(lw.synth_blob_host(k) for k in range(n))
bl = [blobs[i * B:(i + 1) * B] for i in range(n)]


def run(tag):
    t0 = time.perf_counter(); c, p, st = lw.commit_and_prove_batch(blobs, n, s); t1 = time.perf_counter()
    assert not any(st)
    ok = lw.verify_blob_kzg_proof_batch(bl, c, p, s); t2 = time.perf_counter()
    bad = list(p); bad[n // 2] = c[0]
    okb = lw.verify_blob_kzg_proof_batch(bl, c, bad, s)
    cc, st2 = lw.blob_to_kzg_commitment_batch(blobs, n, s)
    pp, st3 = lw.compute_blob_kzg_proof_batch(blobs, b"".join(c), n, s)
    print("%s: commit+prove %.1f ms, verify %.1f ms -> %s, corrupted -> %s" % (tag, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ok, okb), flush=True)
    assert ok is True and okb is False and cc == c and pp == p
    return c, p


c1, p1 = run("1 device ")
lw.set_devices(list(range(g)))
assert lw.get_devices() == list(range(g))
run("%d devices (first call builds the replicas)" % g)
c2, p2 = run("%d devices" % g)
assert c1 == c2 and p1 == p2, "multi-device outputs differ from the single-device outputs"
lw.set_devices([])
c3, p3 = run("1 device again")
assert c3 == c1 and p3 == p1
print("MULTI_DEVICE_CHECK_OK", flush=True)
