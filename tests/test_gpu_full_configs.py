"""GPU tier at the configurations bench.py measures (BASELINE.json configs 2-5), through the C ABI.

Everything here runs with the fixed-base window the bench line is quoted on (the largest that fits HBM), at the full
batch sizes: bit-equality with the oracle on sampled items plus the size-independent properties the domain offers
(whole batch verifies, one replaced proof flips it, tau-oracle for the variable-base MSM).

  * config 2: 1024 synthetic blobs, commitment + blob proof, both MSM kernels           lib.rs:253-283, 361-404
  * config 3: verify_blob_kzg_proof_batch over 4096 blobs, proof #2047 replaced by G    lib.rs:525-692
  * config 5: variable-base MSM at 2^20 and 2^22 against the O(N) tau-oracle            lib.rs:241-243
  * config 4: shard outputs byte-equal to the single-GPU outputs, distributed == monolithic verification
              (needs >= 2 GPUs; torchrun over NCCL)                                     SURVEY §8e
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.py import bls, kzg  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
R = bls.R
GEN = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
BENCH_WINDOW = int(os.environ.get("LWKZG_TEST_BENCH_WINDOW", "16"))   # what bench.py asks for


@pytest.fixture(scope="module")
def lw():
    import lambdaworks_kzg_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def settings_bench(lw):
    """The benchmarked configuration (largest fixed-base table that fits)."""
    old = lw.get_option("window_bits")
    lw.set_option("window_bits", BENCH_WINDOW)
    s = lw.load_trusted_setup_file(os.path.join(GOLDEN, "trusted_setup.txt"))
    yield s
    s.free()
    lw.set_option("window_bits", old)


@pytest.fixture(scope="module")
def ref(py_setup):
    return kzg.RefMode(py_setup)


def _edge_blobs():
    from tests.test_gpu_parity import edge_blobs

    return edge_blobs()


# ------------------------------------------------------------------ config 2
def test_bench_window_is_what_bench_uses(lw, settings_bench):
    # on a 180 GB part the requested window must not have been shrunk
    assert lw.window_bits(settings_bench) == BENCH_WINDOW


@pytest.mark.parametrize("algo", [0, 1])
def test_edge_blobs_at_bench_window(lw, ref, settings_bench, algo):
    blobs = _edge_blobs()
    names = list(blobs)
    old_min = lw.get_option("msm_ba_min_blobs")
    lw.set_option("msm_algo", algo)
    lw.set_option("msm_ba_min_blobs", 1)
    try:
        coms, proofs, st = lw.commit_and_prove_batch(b"".join(blobs[n] for n in names), len(names), settings_bench)
    finally:
        lw.set_option("msm_algo", 1)
        lw.set_option("msm_ba_min_blobs", old_min)
    assert st == [0] * len(names)
    for n, c, p in zip(names, coms, proofs):
        want = ref.blob_to_kzg_commitment(blobs[n])
        assert c == want, n
        assert p == ref.compute_blob_kzg_proof(blobs[n], want), n


def test_1024_blob_batch_at_bench_window_both_kernels(lw, ref, settings_bench):
    """The exact batch bench.py times: 1024 synthetic blobs through the device-pointer entry point; 40 sampled
    blobs against the Python oracle; the XYZZ-only run gives the same bytes; the whole batch verifies."""
    import torch

    n = 1024
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    blobs = torch.empty(n * kzg.BYTES_PER_BLOB, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(blobs.data_ptr(), 0, n, st)
    out = {}
    for algo in (1, 0):
        lw.set_option("msm_algo", algo)
        coms = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
        proofs = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
        status = torch.ones(n, dtype=torch.int32, device=dev)
        lw.commit_and_prove_batch_device(coms.data_ptr(), proofs.data_ptr(), blobs.data_ptr(), n, settings_bench, st, status.data_ptr())
        torch.cuda.synchronize()
        assert int(status.abs().sum()) == 0
        out[algo] = (bytes(coms.cpu().numpy().tobytes()), bytes(proofs.cpu().numpy().tobytes()))
    lw.set_option("msm_algo", 1)
    assert out[0] == out[1]
    cb, pb = out[1]
    sample = sorted(set(range(0, n, 27)) | {1, 255, 256, 511, 512, 1023})
    assert len(sample) >= 32
    for k in sample:
        blob = lw.synth_blob_host(k)
        c = ref.blob_to_kzg_commitment(blob)
        assert cb[48 * k: 48 * k + 48] == c, k
        assert pb[48 * k: 48 * k + 48] == ref.compute_blob_kzg_proof(blob, c), k
    hb = blobs.cpu().numpy().tobytes()
    B = kzg.BYTES_PER_BLOB
    assert lw.verify_blob_kzg_proof_batch([hb[i * B:(i + 1) * B] for i in range(n)], [cb[48 * i:48 * i + 48] for i in range(n)],
                                          [pb[48 * i:48 * i + 48] for i in range(n)], settings_bench) is True


# ------------------------------------------------------------------ config 3
def test_verify_batch_4096_with_proof_2047_replaced(lw, ref, settings_bench):
    """4096 blobs, commitments / proofs from our own commit+prove (spot-checked against the
    oracle), all-valid -> true; proof #2047 replaced by the generator -> false; the batch challenge r equals
    hashlib over the 4096 tuples (utils.rs:166-206)."""
    import hashlib

    import torch

    n = 4096
    B = kzg.BYTES_PER_BLOB
    dev = torch.device("cuda", 0)
    d = torch.empty(n * B, dtype=torch.uint8, device=dev)
    lw.synth_blobs_device(d.data_ptr(), 20000, n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    h_blobs = d.cpu().pin_memory()
    del d
    coms = torch.zeros(n * 48, dtype=torch.uint8).pin_memory()
    proofs = torch.zeros(n * 48, dtype=torch.uint8).pin_memory()
    st = lw.commit_and_prove_batch(h_blobs.data_ptr(), n, settings_bench, coms.data_ptr(), proofs.data_ptr())
    assert not any(st)
    for k in (0, 2047, 4095):
        blob = lw.synth_blob_host(20000 + k)
        assert bytes(h_blobs[k * B:(k + 1) * B].numpy().tobytes()) == blob
        c = ref.blob_to_kzg_commitment(blob)
        assert bytes(coms[48 * k:48 * k + 48].numpy().tobytes()) == c
        assert bytes(proofs[48 * k:48 * k + 48].numpy().tobytes()) == ref.compute_blob_kzg_proof(blob, c)
    assert lw.verify_blob_kzg_proof_batch_ptr(h_blobs.data_ptr(), coms.data_ptr(), proofs.data_ptr(), n, settings_bench) is True
    r_gpu = lw.debug_batch_challenge(settings_bench)
    # tuples (C, z, y, pi): z and y from the oracle would take minutes for 4096 blobs; the phase-1 tuples are
    # checked against the oracle on a sample and hashed on the host
    tuples = lw.verify_batch_phase1(h_blobs.data_ptr(), coms.data_ptr(), proofs.data_ptr(), n, settings_bench)
    for k in (0, 2047, 4095):
        blob = lw.synth_blob_host(20000 + k)
        c = bytes(coms[48 * k:48 * k + 48].numpy().tobytes())
        z = ref.compute_challenge(blob, bls.g1_decompress(c))
        y = kzg.horner(ref.blob_to_coeffs(blob), z)
        t = tuples[160 * k: 160 * k + 160]
        assert t[:48] == c and t[48:80] == z.to_bytes(32, "big") and t[80:112] == y.to_bytes(32, "big")
        assert t[112:] == bytes(proofs[48 * k:48 * k + 48].numpy().tobytes())
    msg = b"RCKZGBATCH___V1_" + (4096).to_bytes(8, "little") + n.to_bytes(8, "little") + tuples
    assert r_gpu == int.from_bytes(hashlib.sha256(msg).digest(), "big") % R
    bad = proofs.clone().pin_memory()
    bad[48 * 2047: 48 * 2048] = torch.frombuffer(bytearray(GEN), dtype=torch.uint8)
    assert lw.verify_blob_kzg_proof_batch_ptr(h_blobs.data_ptr(), coms.data_ptr(), bad.data_ptr(), n, settings_bench) is False
    # device-resident entry point: same answers without the host round trip
    if hasattr(lw, "verify_blob_kzg_proof_batch_device"):
        db, dc, dp, dbad = h_blobs.to(dev), coms.to(dev), proofs.to(dev), bad.to(dev)
        assert lw.verify_blob_kzg_proof_batch_device(db.data_ptr(), dc.data_ptr(), dp.data_ptr(), n, settings_bench) is True
        assert lw.verify_blob_kzg_proof_batch_device(db.data_ptr(), dc.data_ptr(), dbad.data_ptr(), n, settings_bench) is False


# ------------------------------------------------------------------ config 5
def _splitmix_scalars(seed, n):
    """numpy restatement of the synthetic scalar stream (csrc/misc.cu): four SplitMix64 outputs per item,
    big-endian, top two bits cleared."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    t = np.arange(n, dtype=np.uint64)
    st = np.uint64(0xB2004844) ^ ((np.uint64(seed) * np.uint64(4096) + t) & M)
    words = []
    with np.errstate(over="ignore"):
        for _ in range(4):
            st = st + np.uint64(0x9E3779B97F4A7C15)
            z = st.copy()
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            words.append(z ^ (z >> np.uint64(31)))
    words[0] = words[0] & np.uint64(0x3FFFFFFFFFFFFFFF)
    return words  # most significant first


@pytest.mark.parametrize("lg", [20, 22])
def test_var_msm_full_sizes_vs_tau_oracle(lw, settings_bench, py_setup, lg):
    n = 1 << lg
    seed = 9
    ms, got = lw.bench_var_msm(n, settings_bench, iters=1, seed=seed)
    dlog = lw.synth_point_dlogs(settings_bench, n, py_setup.tau)          # discrete logs of the synthetic points
    w = _splitmix_scalars(seed, n)
    acc = 0
    w0, w1, w2, w3 = (x.tolist() for x in w)
    for t in range(n):
        k = (w0[t] << 192) | (w1[t] << 128) | (w2[t] << 64) | w3[t]
        acc += k * dlog[t]
        if (t & 0xFFFF) == 0:
            acc %= R
    assert got == bls.g1_compress(bls.g1_mul(bls.G1, acc % R)), lg


# ------------------------------------------------------------------ config 4 / SURVEY §8e
def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_byte_equality(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, NB="2048", WB="13")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "MULTI_GPU_CHECK_OK" in p.stdout


def test_in_library_multi_device_byte_equality(lw, ref):
    """lwkzg_set_devices: one C call drives every visible GPU; outputs are byte-equal to the single-device call."""
    if _gpu_count() < 2 or not hasattr(lw, "set_devices"):
        pytest.skip("needs >= 2 GPUs")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_device_check.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "MULTI_DEVICE_CHECK_OK" in p.stdout
