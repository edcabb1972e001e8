#!/usr/bin/env python3
"""lwkzg_set_devices for the PeerDAS calls: ONE process, one C call, every visible GPU.  The sharded cell batch must
return the bytes of the single-device call (MODE_DENEB: the replicas get the monomial SRS from the primary context).
Prints MULTI_DEVICE_CELLS_OK on success."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

g = torch.cuda.device_count()
assert g >= 2, "needs at least two GPUs"
n = int(os.environ.get("NB", "512"))
lw.set_option("mode", 2)
lw.set_option("window_bits", 10)
lw.set_option("cell_window_bits", int(os.environ.get("CWB", "12")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
blobs = b"".join(lw.synth_blob_host(k) for k in range(n))


def run(tag):
    t0 = time.perf_counter()
    cells, proofs, st = lw.compute_cells_and_kzg_proofs_batch(blobs, n, s)
    dt = time.perf_counter() - t0
    assert not any(st)
    print("%s: %d blobs in %.1f ms (%.0f blobs/s)" % (tag, n, dt * 1e3, n / dt), flush=True)
    return cells, proofs


run("1 device (first call builds the FK20 tables)")
c1, p1 = run("1 device ")
lw.set_devices(list(range(g)))
run("%d devices (first call builds the replicas)" % g)
c2, p2 = run("%d devices" % g)
assert c1 == c2 and p1 == p2, "multi-device cell outputs differ from the single-device outputs"
lw.set_devices([])
print("MULTI_DEVICE_CELLS_OK", flush=True)
