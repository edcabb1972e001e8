"""Device-resident cells + proofs for a few (batch, cell_chunk_blobs) pairs: python tools/cells_chunk_sweep.py n:chunk ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lambdaworks_kzg_b200 as lw

lw.set_option("mode", 2)
lw.set_option("window_bits", 8)
s = lw.load_trusted_setup_file(os.path.join(ROOT, "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
for spec in sys.argv[1:]:
    n, chunk = (int(x) for x in spec.split(":"))
    lw.set_option("cell_chunk_blobs", chunk)
    d_blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(d_blobs.data_ptr(), 3 << 20, n, stream)
    d_proofs = torch.zeros(n * 128 * 48, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    for _ in range(2):
        lw.compute_cells_and_kzg_proofs_batch_device(0, d_proofs.data_ptr(), d_blobs.data_ptr(), n, s, stream, d_st.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lw.compute_cells_and_kzg_proofs_batch_device(0, d_proofs.data_ptr(), d_blobs.data_ptr(), n, s, stream, d_st.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("n=%d chunk=%d: %.2f ms  %.0f blobs/s" % (n, chunk, ms, n / ms * 1e3), flush=True)
    del d_blobs, d_proofs
s.free()
