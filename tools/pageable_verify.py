"""4096-blob verification from pageable host memory for a few staging-thread counts (LWKZG_STAGE_THREADS is read once per process)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

n = 4096
lw.set_option("window_bits", 13)
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d_blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(d_blobs.data_ptr(), 0, n, st)
d_c = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_p = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
torch.cuda.synchronize()
hb, hc, hp = d_blobs.cpu(), d_c.cpu(), d_p.cpu()
fn = lambda: lw.verify_blob_kzg_proof_batch_ptr(hb.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, s)
fn()
ts = []
for _ in range(7):
    t = time.perf_counter(); ok = fn(); ts.append((time.perf_counter() - t) * 1e3)
print("threads=%s cores=%d: pageable verify min %.2f median %.2f ms -> %s" % (os.environ.get("LWKZG_STAGE_THREADS", "default"), os.cpu_count(), min(ts), sorted(ts)[3], ok), flush=True)
