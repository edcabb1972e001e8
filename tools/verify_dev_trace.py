"""Device-resident 4096-blob verification with LWKZG_VERIFY_TRACE=1: where the time goes when no copy paces the chunks."""
import os, sys, time
if os.environ.get("TRACE", "0") == "1":
    os.environ["LWKZG_VERIFY_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lambdaworks_kzg_b200 as lw

if os.environ.get("CACHECFG"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.cuda.init()
    print("cudaDeviceSetCacheConfig ->", rt.cudaDeviceSetCacheConfig(int(os.environ["CACHECFG"])), flush=True)
n = int(os.environ.get("NB", "4096"))
lw.set_option("window_bits", int(os.environ.get("WB", "13")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", 0)
d_blobs = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
lw.synth_blobs_device(d_blobs.data_ptr(), 0, n, st)
d_c = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_p = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
lw.commit_and_prove_batch_device(d_c.data_ptr(), d_p.data_ptr(), d_blobs.data_ptr(), n, s, st, d_st.data_ptr())
torch.cuda.synchronize()
hb, hc, hp = d_blobs.cpu().pin_memory(), d_c.cpu().pin_memory(), d_p.cpu().pin_memory()
for ns in [int(x) for x in os.environ.get("STREAMS", "6").split(",")]:
  for split in [int(x) for x in os.environ.get("SPLIT", "1").split(",")]:
    lw.set_option("verify_streams", ns)
    lw.set_option("verify_split_subgroup", split)
    lw.set_option("verify_overlap_decode", int(os.environ.get("OVERLAP", "1")))
    for name, fn in (("device-resident", lambda: lw.verify_blob_kzg_proof_batch_device(d_blobs.data_ptr(), d_c.data_ptr(), d_p.data_ptr(), n, s)),
                     ("pinned", lambda: lw.verify_blob_kzg_proof_batch_ptr(hb.data_ptr(), hc.data_ptr(), hp.data_ptr(), n, s))):
        fn()
        ts = []
        for rep in range(5):
            t = time.perf_counter()
            ok = fn()
            ts.append((time.perf_counter() - t) * 1e3)
        print("split=%d streams=%d %s verify n=%d -> %s: min %.2f ms, median %.2f ms" % (split, ns, name, n, ok, min(ts), sorted(ts)[2]), flush=True)
