#!/usr/bin/env python3
"""One blob-proof batch and one batched verification, for a per-kernel launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/stage_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw
from lambdaworks_kzg_b200 import api

n = int(os.environ.get("NB", "1024"))
lw.set_option("window_bits", int(os.environ.get("WB", "10")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
lib = api.load_library(); sp = api._sp(s)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(blobs.data_ptr(), 0, n, st)
coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
status = torch.zeros(n, dtype=torch.int32, device=dev)
lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, s, st, status.data_ptr())
torch.cuda.synchronize()
lib.lwkzg_compute_blob_kzg_proof_batch_device(proofs.data_ptr(), blobs.data_ptr(), coms.data_ptr(), n, sp, st, status.data_ptr())
torch.cuda.synchronize()
hb = blobs.cpu().pin_memory(); hc = coms.cpu().pin_memory(); hp = proofs.cpu().pin_memory()
ok = lw.verify_blob_kzg_proof_batch_ptr(hb.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, s)
print("verify:", ok)
s.free()
