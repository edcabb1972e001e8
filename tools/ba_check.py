#!/usr/bin/env python3
"""GPU check of the batched-affine MSM kernel: identical commitments/proofs to the XYZZ kernel on the
same blobs (synthetic + degenerate ones), and the kernel-only time of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

c = int(os.environ.get("WB", "16"))
n = int(os.environ.get("NB", "1024"))
lw.set_option("window_bits", c)
t = time.time()
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
torch.cuda.synchronize()
print("setup load + table build (c=%d): %.2f s" % (c, time.time() - t), flush=True)
dev = torch.device("cuda", 0)
blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
lw.synth_blobs_device(blobs.data_ptr(), 0, n, st)
# degenerate blobs: all zero, all words equal, single non-zero word, words = r - 1
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
v = blobs.view(n, 4096, 32)
if os.environ.get("DEGEN", "1") == "1":
  if True:
    v[1].zero_()
    v[2][:] = torch.tensor(list((12345678901234567890123456789).to_bytes(32, "big")), dtype=torch.uint8, device=dev)
    v[3].zero_(); v[3][4095][31] = 1
    v[4][:] = torch.tensor(list((R - 1).to_bytes(32, "big")), dtype=torch.uint8, device=dev)
    v[5][:] = torch.tensor(list((1).to_bytes(32, "big")), dtype=torch.uint8, device=dev)
    v[6][:] = torch.tensor(list((R + 5).to_bytes(32, "big")), dtype=torch.uint8, device=dev)
res = {}
def run_cp(tag):
    coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
    best = 1e9
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, s, st, 0)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s commit+prove n=%d: %.2f ms -> %.1f blobs/s" % (tag, n, best, n / best * 1e3), flush=True)
    return coms.cpu(), proofs.cpu()

for algo in [int(x) for x in os.environ.get("ALGOS", "0,1").split(",")]:
    lw.set_option("msm_algo", algo)
    if algo == 0:
        res[0] = run_cp("algo=0")
        for nb in sorted({min(512, n), n}):
            ms = lw.bench_msm_kernel(blobs.data_ptr(), nb, s, 2, 3)
            print("algo=0 msm kernel n=%d bpb=2: %.3f ms -> %.1f MSM/s" % (nb, ms, nb / ms * 1e3), flush=True)
        continue
    for variant in [int(x) for x in os.environ.get("VARIANTS", "0,1,2,3").split(",")]:
        lw.set_option("msm_ba_variant", variant)
        for nb in sorted({min(512, n), min(1024, n), n}):
            ms = lw.bench_msm_kernel(blobs.data_ptr(), nb, s, 0, 3)
            print("algo=1 variant=%d msm kernel n=%d: %.3f ms -> %.1f MSM/s" % (variant, nb, ms, nb / ms * 1e3), flush=True)
        for chunk in (512, 1024):
            lw.set_option("chunk_blobs", chunk)
            r = run_cp("algo=1 variant=%d chunk=%d" % (variant, chunk))
            if 1 in res and not (torch.equal(r[0], res[1][0]) and torch.equal(r[1], res[1][1])):
                print("VARIANT MISMATCH", variant, chunk); sys.exit(1)
            res[1] = r
        lw.set_option("chunk_blobs", 512)
if len(res) < 2:
    sys.exit(0)
same_c = torch.equal(res[0][0], res[1][0]); same_p = torch.equal(res[0][1], res[1][1])
print("commitments identical:", same_c, " proofs identical:", same_p)
if not (same_c and same_p):
    a, b = res[0][0].view(n, 48), res[1][0].view(n, 48)
    bad = [i for i in range(n) if not torch.equal(a[i], b[i])]
    print("differing commitments:", bad[:20], len(bad))
    a, b = res[0][1].view(n, 48), res[1][1].view(n, 48)
    bad = [i for i in range(n) if not torch.equal(a[i], b[i])]
    print("differing proofs:", bad[:20], len(bad))
    sys.exit(1)
print("commit[0]", bytes(res[1][0][:48].numpy().tobytes()).hex())
