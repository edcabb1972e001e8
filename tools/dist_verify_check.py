#!/usr/bin/env python3
"""Multi-GPU batched verification over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_verify_check.py

Every rank generates its contiguous shard of NB synthetic blobs, commits and proves it (no communication), then the
ranks run verify_blob_kzg_proof_batch_distributed (two tiny all-gathers).  Rank 0 checks the result against the
single-GPU monolithic call on the full batch, and a corrupted proof must flip it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import lambdaworks_kzg_b200 as lw

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(os.environ.get("NB", "512"))
lw.set_option("window_bits", int(os.environ.get("WB", "10")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
first, cnt = lw.shard_range(n_total, world, rank)
blobs = b"".join(lw.synth_blob_host(k) for k in range(first, first + cnt))
coms, proofs, st = lw.commit_and_prove_batch(blobs, cnt, s)
assert not any(st)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
ok = lw.verify_blob_kzg_proof_batch_distributed(blobs, b"".join(coms), b"".join(proofs), n_total, s)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
bad_proofs = list(proofs)
if rank == world - 1:
    bad_proofs[-1] = coms[0]
ok_bad = lw.verify_blob_kzg_proof_batch_distributed(blobs, b"".join(coms), b"".join(bad_proofs), n_total, s)
# gather everything on rank 0 for the monolithic cross-check
outs = [None] * world
dist.all_gather_object(outs, (coms, proofs))
if rank == 0:
    all_blobs = [lw.synth_blob_host(k) for k in range(n_total)]
    all_c = [c for o in outs for c in o[0]]
    all_p = [p for o in outs for p in o[1]]
    mono = lw.verify_blob_kzg_proof_batch(all_blobs, all_c, all_p, s)
    print("world=%d n=%d distributed=%s monolithic=%s corrupted=%s  %.1f ms" % (world, n_total, ok, mono, ok_bad, dt * 1e3), flush=True)
    assert ok is True and mono is True and ok_bad is False
dist.destroy_process_group()
