"""CPU tier: the N > 1 path.  Blob sharding and the two tiny exchanges of the
batched verification, run with world_size = 2 over gloo; the per-rank compute
phases are served by the Python oracle here (the GPU tier runs the same driver
over the C ABI)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.py import bls, kzg  # noqa: E402
from lambdaworks_kzg_b200.sharding import shard_range, verify_blob_kzg_proof_batch_distributed  # noqa: E402
from lambdaworks_kzg_b200 import api  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 8, 1024, 262144, 262145):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                first, cnt = shard_range(n, world, r)
                covered += list(range(first, first + cnt)) if n < 2000 else []
                assert 0 <= cnt <= (n + world - 1) // world
            if n < 2000:
                assert covered == list(range(n))
            assert sum(shard_range(n, world, r)[1] for r in range(world)) == n


class OraclePhases:
    """The three phases of lwkzg_verify_batch_phase{1,2,3}, restated with oracle/py."""

    def __init__(self, setup):
        self.ref = kzg.RefMode(setup)
        self.local = None

    def single(self, blob, commitment, proof):
        try:
            return self.ref.verify_blob_kzg_proof(blob, commitment, proof)
        except kzg.KzgError as e:
            raise api.KzgError(e.code, "oracle")

    def phase1(self, blobs, commitments, proofs, n_local):
        out, items = b"", []
        B = kzg.BYTES_PER_BLOB
        try:
            for i in range(n_local):
                c = self.ref._decompress(commitments[48 * i: 48 * i + 48])
                blob = blobs[B * i: B * i + B]
                z = self.ref.compute_challenge(blob, c)
                y = kzg.horner(self.ref.blob_to_coeffs(blob), z)
                pi = self.ref._decompress(proofs[48 * i: 48 * i + 48])
                items.append((c, z, y, pi))
                out += bls.g1_compress(c) + z.to_bytes(32, "big") + y.to_bytes(32, "big") + bls.g1_compress(pi)
        except kzg.KzgError as e:
            raise api.KzgError(e.code, "oracle")
        self.local = items
        return out

    def phase2(self, all_tuples, n_total, first, n_local):
        import hashlib

        msg = kzg.RANDOM_CHALLENGE_KZG_BATCH_DOMAIN + (4096).to_bytes(8, "little") + n_total.to_bytes(8, "little") + all_tuples
        r = int.from_bytes(hashlib.sha256(msg).digest(), "big") % bls.R
        a = b = c = None
        ysum = 0
        for k, (cm, z, y, pi) in enumerate(self.local):
            ri = pow(r, first + k, bls.R)
            a = bls.g1_add(a, bls.g1_mul(pi, ri))
            b = bls.g1_add(b, bls.g1_mul(pi, ri * z % bls.R))
            c = bls.g1_add(c, bls.g1_mul(cm, ri))
            ysum = (ysum + ri * y) % bls.R
        c = bls.g1_add(c, bls.g1_neg(bls.g1_mul(bls.G1, ysum)))
        enc = lambda p: bytes(96) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")  # noqa: E731
        return enc(a) + enc(b) + enc(c)

    def phase3(self, partials, n_ranks):
        dec = lambda b: None if not any(b) else (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big"))  # noqa: E731
        pl = pz = cy = None
        for r in range(n_ranks):
            base = partials[288 * r: 288 * r + 288]
            pl = bls.g1_add(pl, dec(base[:96]))
            pz = bls.g1_add(pz, dec(base[96:192]))
            cy = bls.g1_add(cy, dec(base[192:]))
        return self.ref._pairing_check(bls.g1_add(cy, pz), pl)


def _worker(rank, world, port, n_total, corrupt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        setup = kzg.parse_setup_text(open(os.path.join(ROOT, "tests", "golden", "trusted_setup.txt")).read())
        ref = kzg.RefMode(setup)
        blobs, coms, proofs = [], [], []
        for k in range(n_total):
            coeffs = [(k + 1) * 1000003 + i * i for i in range(6 + k)]
            blob = b"".join(c.to_bytes(32, "big") for c in coeffs).ljust(kzg.BYTES_PER_BLOB, b"\0")
            c = ref.blob_to_kzg_commitment(blob)
            blobs.append(blob)
            coms.append(c)
            proofs.append(ref.compute_blob_kzg_proof(blob, c))
        if corrupt == "proof":
            proofs[n_total - 1] = coms[0]
        if corrupt == "invalid":
            coms[0] = bytes(48)
        first, cnt = shard_range(n_total, world, rank)
        phases = OraclePhases(setup)
        try:
            got = verify_blob_kzg_proof_batch_distributed(b"".join(blobs[first:first + cnt]), b"".join(coms[first:first + cnt]),
                                                          b"".join(proofs[first:first + cnt]), n_total, None, dist=dist,
                                                          device=torch.device("cpu"), phases=phases)
        except api.KzgError as e:
            got = "error%d" % e.code
        try:
            want = ref.verify_blob_kzg_proof_batch(blobs, coms, proofs)
        except kzg.KzgError as e:
            want = "error%d" % e.code
        q.put((rank, got, want))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_total,corrupt", [(5, None), (5, "proof"), (3, "invalid"), (1, None), (0, None)])
def test_distributed_batch_verify_gloo_world2(n_total, corrupt):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, corrupt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, want in res:
        assert got == want, (rank, got, want)
    assert len({g for _, g, _ in res}) == 1  # same answer on every rank
