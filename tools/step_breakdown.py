#!/usr/bin/env python3
"""Where a commit+proof step spends its time: isolated MSM kernel vs the chunked device-API drivers
(commit only / blob proof only / commit+proof) for several chunk sizes.  CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw
from lambdaworks_kzg_b200 import api

lw.set_option("window_bits", int(os.environ.get("WB", "15")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
lib = api.load_library()
sp = api._sp(s)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("REPS", "4"))


def timed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in [int(x) for x in os.environ.get("NB", "1024,3552").split(",")]:
    blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(blobs.data_ptr(), 0, n, stream)
    coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    k_ms = lw.bench_msm_kernel(blobs.data_ptr(), n, s, 0, reps)
    print("n=%d isolated MSM launch: %.2f ms = %.1f K MSM/s" % (n, k_ms, n / k_ms), flush=True)
    for chunk in [int(x) for x in os.environ.get("CHUNKS", "256,512,100000").split(",")]:
        lw.set_option("chunk_blobs", chunk)
        t_c = timed(lambda: lib.lwkzg_blob_to_kzg_commitment_batch_device(coms.data_ptr(), blobs.data_ptr(), n, sp, stream))
        t_p = timed(lambda: lib.lwkzg_compute_blob_kzg_proof_batch_device(proofs.data_ptr(), blobs.data_ptr(), coms.data_ptr(), n, sp, stream, status.data_ptr()))
        t_cp = timed(lambda: lib.lwkzg_commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, sp, stream, status.data_ptr()))
        print("  chunk %6d: commit %.2f ms (%.1f K MSM/s)  blob proof %.2f ms (%.1f K MSM/s)  commit+proof %.2f ms (%.1f K MSM/s, %.1f K blobs/s)" % (
            chunk, t_c, n / t_c, t_p, n / t_p, t_cp, 2 * n / t_cp, n / t_cp), flush=True)
    lw.set_option("chunk_blobs", 256)
    # steps kept in flight on alternating user streams (the driver no longer waits on the host between calls)
    for depth in (2, 3):
        side = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        outs = [(torch.zeros(n * 48, dtype=torch.uint8, device=dev), torch.zeros(n * 48, dtype=torch.uint8, device=dev),
                 torch.zeros(n, dtype=torch.int32, device=dev)) for _ in range(depth)]
        main = torch.cuda.current_stream()

        def run(k=8):
            for sd in side:
                sd.wait_stream(main)
            for i in range(k):
                c, p, st = outs[i % depth]
                lib.lwkzg_commit_and_prove_batch_device(c.data_ptr(), p.data_ptr(), blobs.data_ptr(), n, sp, side[i % depth].cuda_stream, st.data_ptr())
            for sd in side:
                main.wait_stream(sd)
        run(2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(8); e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 8
        ok = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs) and torch.equal(outs[0][0], coms)
        print("  %d steps in flight (chunk 256): %.2f ms per step (%.1f K blobs/s)  outputs equal: %s" % (depth, t, n / t, ok), flush=True)
    del blobs
s.free()
