#!/usr/bin/env python3
"""One process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py

* every rank commits to and proves its contiguous shard of NB synthetic blobs (no communication); rank 0 recomputes the
  whole batch alone and the shard outputs must be byte-equal to it;
* the batch is verified three ways -- monolithic on rank 0, distributed with host exchanges (gloo-style bytes path) and
  distributed with every exchanged byte in device memory -- and all must agree, also after corrupting one proof.
Prints MULTI_GPU_CHECK_OK on success."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import lambdaworks_kzg_b200 as lw

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(os.environ.get("NB", "2048"))
lw.set_option("window_bits", int(os.environ.get("WB", "13")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
dev = torch.device("cuda", local)
B = 131072
first, cnt = lw.shard_range(n_total, world, rank)
d_blobs = torch.empty(cnt * B, dtype=torch.uint8, device=dev)
lw.synth_blobs_device(d_blobs.data_ptr(), first, cnt, 0)
torch.cuda.synchronize()
blobs = bytes(d_blobs.cpu().numpy().tobytes())
coms, proofs, st = lw.commit_and_prove_batch(blobs, cnt, s)
assert not any(st)
outs = [None] * world
dist.all_gather_object(outs, (coms, proofs))
all_c = [c for o in outs for c in o[0]]
all_p = [p for o in outs for p in o[1]]
if rank == 0:
    full = torch.empty(n_total * B, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(full.data_ptr(), 0, n_total, 0)
    torch.cuda.synchronize()
    full_b = bytes(full.cpu().numpy().tobytes())
    c1, p1, st1 = lw.commit_and_prove_batch(full_b, n_total, s)
    assert not any(st1)
    assert c1 == all_c and p1 == all_p, "shard outputs differ from the single-GPU outputs"
    mono = lw.verify_blob_kzg_proof_batch([full_b[i * B:(i + 1) * B] for i in range(n_total)], all_c, all_p, s)
    assert mono is True
    del full
dist.barrier()
cb, pb = b"".join(coms), b"".join(proofs)
ok_host = lw.verify_blob_kzg_proof_batch_distributed(blobs, cb, pb, n_total, s)
d_c = torch.frombuffer(bytearray(cb), dtype=torch.uint8).to(dev)
d_p = torch.frombuffer(bytearray(pb), dtype=torch.uint8).to(dev)
torch.cuda.synchronize(); dist.barrier()
times = []
for rep in range(3):
    t0 = time.perf_counter()
    ok_dev = lw.verify_blob_kzg_proof_batch_distributed_device(d_blobs.data_ptr(), d_c.data_ptr(), d_p.data_ptr(), n_total, s, inputs_on_device=True)
    torch.cuda.synchronize(); dist.barrier()
    times.append((time.perf_counter() - t0) * 1e3)
bad = list(proofs)
if rank == world - 1:
    bad[-1] = coms[0]
d_bad = torch.frombuffer(bytearray(b"".join(bad)), dtype=torch.uint8).to(dev)
ok_bad = lw.verify_blob_kzg_proof_batch_distributed_device(d_blobs.data_ptr(), d_c.data_ptr(), d_bad.data_ptr(), n_total, s, inputs_on_device=True)
ok_bad_host = lw.verify_blob_kzg_proof_batch_distributed(blobs, cb, b"".join(bad), n_total, s)
assert ok_host is True and ok_dev is True and ok_bad is False and ok_bad_host is False, (ok_host, ok_dev, ok_bad, ok_bad_host)
if rank == 0:
    print("world=%d n=%d: shard outputs == single GPU, distributed verify (host / device exchanges) == monolithic; "
          "device-resident distributed verify %.1f ms (best of 3)" % (world, n_total, min(times)), flush=True)
    print("MULTI_GPU_CHECK_OK", flush=True)
dist.destroy_process_group()
