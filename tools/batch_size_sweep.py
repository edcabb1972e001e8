"""commit+proof throughput against batch size (device-resident), with the batched-affine kernel forced on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

lw.set_option("window_bits", int(os.environ.get("WB", "16")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
N = 2048
d_blobs = torch.empty(N * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(d_blobs.data_ptr(), 0, N, st)
d_c = torch.zeros(N * 48, dtype=torch.uint8, device=dev)
d_p = torch.zeros(N * 48, dtype=torch.uint8, device=dev)
d_st = torch.zeros(N, dtype=torch.int32, device=dev)
for n in (1, 2, 3, 4, 6, 8, 16, 32, 64, 128, 192, 256, 384, 512, 1024, 2048):
    row = []
    for ba_min in (1, 1 << 20):
        lw.set_option("msm_ba_min_blobs", ba_min)
        for _ in range(2):
            lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
        torch.cuda.synchronize()
        reps = 3 if n >= 256 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row.append((ms, n / ms * 1e3))
    print("n=%4d  batched-affine: %7.2f ms %7.0f blobs/s   XYZZ: %7.2f ms %7.0f blobs/s" % (n, row[0][0], row[0][1], row[1][0], row[1][1]), flush=True)
lw.set_option("msm_ba_min_blobs", 5)
