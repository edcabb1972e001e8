#!/usr/bin/env python3
"""Runs the dominant kernel alone (for ncu): NB synthetic blobs, ITERS launches of the fixed-base MSM kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

n = int(os.environ.get("NB", "1024"))
lw.set_option("window_bits", int(os.environ.get("WB", "16")))
lw.set_option("msm_ba_variant", int(os.environ.get("VARIANT", "0")))
lw.set_option("msm_algo", int(os.environ.get("ALGO", "1")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(blobs.data_ptr(), 0, n, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ms = lw.bench_msm_kernel(blobs.data_ptr(), n, s, int(os.environ.get("BPB", "0")), int(os.environ.get("ITERS", "2")))
print("n=%d variant=%s: %.3f ms per launch -> %.1f MSM/s" % (n, os.environ.get("VARIANT", "0"), ms, n / ms * 1e3))
s.free()
