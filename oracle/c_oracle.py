"""ctypes wrapper of the C restatement oracle (oracle/c/kzg_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the shipped package.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "build", "libkzg_oracle.so")
BLOB = 4096 * 32


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "kzg_ref.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "c")])
    return LIB


class COracle:
    def __init__(self, setup_text: str):
        self.lib = ctypes.CDLL(build())
        lines = setup_text.splitlines()
        n1 = int(lines[0])
        g1 = b"".join(bytes.fromhex(x.strip()) for x in lines[2: 2 + n1])
        rc = self.lib.oracle_load_setup_g1(g1, n1)
        if rc:
            raise RuntimeError("oracle_load_setup_g1 -> %d" % rc)
        self.lib.oracle_g1_values.restype = ctypes.c_void_p

    def g1_values_bytes(self) -> bytes:
        return ctypes.string_at(self.lib.oracle_g1_values(), 4096 * 144)

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    def blob_to_kzg_commitment(self, blob: bytes):
        out = ctypes.create_string_buffer(48)
        rc = self.lib.oracle_blob_to_kzg_commitment(out, blob)
        return rc, out.raw

    def compute_kzg_proof(self, blob: bytes, z: bytes):
        p, y = ctypes.create_string_buffer(48), ctypes.create_string_buffer(32)
        rc = self.lib.oracle_compute_kzg_proof(p, y, blob, z)
        return rc, p.raw, y.raw

    def compute_blob_kzg_proof(self, blob: bytes, commitment: bytes):
        out = ctypes.create_string_buffer(48)
        rc = self.lib.oracle_compute_blob_kzg_proof(out, blob, commitment)
        return rc, out.raw

    def commit_and_prove_batch(self, blobs: bytes, n: int, nthreads: int = 0):
        c, p = ctypes.create_string_buffer(48 * n), ctypes.create_string_buffer(48 * n)
        rc = self.lib.oracle_commit_and_prove_batch(c, p, blobs, n, nthreads)
        return rc, c.raw, p.raw

    def g1_decompress_check(self, b: bytes):
        out = ctypes.create_string_buffer(48)
        ok = self.lib.oracle_g1_decompress_check(b, out)
        return bool(ok), out.raw

    def sha256(self, msg: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.oracle_sha256(out, msg, ctypes.c_size_t(len(msg)))
        return out.raw

    # ---- bench.py helpers
    def synth_blob(self, k: int) -> bytes:
        """SURVEY 8(d) synthetic blob k (same bytes as the GPU generator, csrc/misc.cu)."""
        buf = ctypes.create_string_buffer(BLOB)
        self.lib.oracle_synth_blob(buf, ctypes.c_uint64(k))
        return buf.raw

    def g1_lincomb(self, points_xy_be: bytes, scalars_be: bytes, n: int):
        """g1_lincomb (src/lib.rs:241-243) -> (rc, 48-byte compressed sum); rc 1 = point off the curve."""
        out = ctypes.create_string_buffer(48)
        rc = self.lib.oracle_g1_lincomb(out, points_xy_be, scalars_be, n)
        return rc, out.raw

    def synth_msm_inputs(self, n: int, seed: int = 0):
        pts, sc = ctypes.create_string_buffer(96 * n), ctypes.create_string_buffer(32 * n)
        rc = self.lib.oracle_synth_msm_inputs(pts, sc, n, ctypes.c_uint64(seed))
        if rc:
            raise RuntimeError("oracle_synth_msm_inputs -> %d" % rc)
        return pts.raw, sc.raw
