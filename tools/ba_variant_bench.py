"""Batched-affine MSM kernel variants side by side (device-resident commit+proof step and the kernel alone)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

n = int(os.environ.get("NB", "1024"))
lw.set_option("window_bits", int(os.environ.get("WB", "16")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d_blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(d_blobs.data_ptr(), 0, n, st)
d_c = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_p = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
ref = None
for v in [int(x) for x in os.environ.get("VARIANTS", "0,1").split(",")]:
    lw.set_option("msm_ba_variant", v)
    for _ in range(2):
        lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
    torch.cuda.synchronize()
    out = (bytes(d_c.cpu().numpy()), bytes(d_p.cpu().numpy()))
    if ref is None:
        ref = out
    assert out == ref, "variant %d differs" % v
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    k_ms = lw.bench_msm_kernel(d_blobs.data_ptr(), n, s, 0, 5)
    print("variant %d: step %.2f ms = %.0f blobs/s; kernel alone %.2f ms" % (v, ms, n / ms * 1e3, k_ms), flush=True)
