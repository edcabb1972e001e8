#!/usr/bin/env python3
"""First-light timing on a GPU box: table build, IMAD peak, commit+prove batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

c = int(os.environ.get("WB", "13"))
n = int(os.environ.get("NB", "1024"))
print("device", torch.cuda.get_device_name(0), flush=True)
for v in (0, 1):
    print("imad_peak variant", v, "%.3e MAC32/s" % lw.imad_peak(v), flush=True)
lw.set_option("window_bits", c)
t = time.time()
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
torch.cuda.synchronize()
print("setup load + table build (c=%d): %.2f s" % (c, time.time() - t), flush=True)
print("mem used GB", (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30, flush=True)
dev = torch.device("cuda", 0)
blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
lw.synth_blobs_device(blobs.data_ptr(), 0, n, st)
coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
for bpb in [int(x) for x in os.environ.get("BPB", "0").split(",")]:
    lw.set_option("msm_blocks_per_blob", bpb)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, s, st, 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("bpb=%d commit+prove n=%d: %.2f ms -> %.1f blobs/s" % (bpb, n, ms, n / ms * 1e3), flush=True)
print("commit[0]", bytes(coms[:48].cpu().numpy().tobytes()).hex())
print("proof[0]", bytes(proofs[:48].cpu().numpy().tobytes()).hex())
print("launches", lw.kernel_launches())
