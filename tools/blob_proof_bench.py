"""BASELINE config 2: compute_blob_kzg_proof for 1024 device-resident blobs with given commitments."""
import os, sys
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
d_p2 = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
torch.cuda.synchronize()
for name, fn in (("commit only", lambda: lw.blob_to_kzg_commitment_batch_device(d_c.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())),
                 ("blob proofs for given commitments", lambda: lw.compute_blob_kzg_proof_batch_device(d_p2.data_ptr(), d_blobs.data_ptr(), d_c.data_ptr(), n, s, st, d_st.data_ptr())),
                 ("commit + proof", lambda: lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p2.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr()))):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s, %d blobs: %.2f ms = %.0f blobs/s" % (name, n, ms, n / ms * 1e3), flush=True)
assert bytes(d_p.cpu().numpy()) == bytes(d_p2.cpu().numpy())
