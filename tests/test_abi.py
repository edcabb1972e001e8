"""CPU tier: the C-ABI library loads and exports every symbol include/lwkzg.h
declares (no compute calls -- there is no GPU here), and the product has no CPU
compute path to fall back to."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lwkzg.h")


@pytest.fixture(scope="module")
def lib():
    import lambdaworks_kzg_b200 as lw

    if not os.path.exists(lw.lib_path()):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lambdaworks_kzg_b200", "csrc"), "-j", "8"])
    return lw.load_library()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if not n.startswith("__")))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    # the 9 c-kzg-4844 entry points of the reference (src/lib.rs:253-829)
    for must in ["blob_to_kzg_commitment", "compute_kzg_proof", "compute_blob_kzg_proof", "verify_kzg_proof", "verify_blob_kzg_proof",
                 "verify_blob_kzg_proof_batch", "load_trusted_setup", "load_trusted_setup_file", "free_trusted_setup"]:
        assert must in names
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n


def test_struct_layouts_match_reference():
    from lambdaworks_kzg_b200.api import CKZGSettings

    assert ctypes.sizeof(CKZGSettings) == 24          # 3 pointers, lib.rs:210-222
    # blst_p1 = 3 x 6 x u64 = 144 B, blst_p2 = 288 B (lib.rs:110-159)
    src = open(HEADER).read()
    assert "typedef struct { limb_t l[6]; } blst_fp;" in src   # NOT lambdaworks_kzg.h's wrong l[4]
    assert "C_KZG_OK = 0" in src


def test_no_cpu_fallback_in_product():
    """The shipped package must not import the oracle, and the library has no
    host implementation of the math (host emulation lives under tests/ only)."""
    pkg = os.path.join(ROOT, "lambdaworks_kzg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "c_oracle" not in txt, f
    mk = open(os.path.join(pkg, "csrc", "Makefile")).read()
    assert "LWKZG_HOST_EMUL" not in mk
    assert "compute_100a" in mk and "sm_100a" in mk


def test_compute_call_fails_loudly_without_gpu(lib):
    import lambdaworks_kzg_b200 as lw

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(lw.KzgError):
        lw.load_trusted_setup_file(os.path.join(ROOT, "tests", "golden", "trusted_setup.txt"))


def test_synth_blob_generator_spec():
    """SplitMix64 seeded with 0xB2004844 ^ (k*4096+i), 4 BE u64, byte0 &= 0x3f."""
    import lambdaworks_kzg_b200 as lw

    M = (1 << 64) - 1

    def sm(state):
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return state, z ^ (z >> 31)

    for k in (0, 3, 262143):
        blob = lw.synth_blob_host(k)
        for i in (0, 1, 4095):
            st = 0xB2004844 ^ (k * 4096 + i)
            w = b""
            for _ in range(4):
                st, v = sm(st)
                w += v.to_bytes(8, "big")
            w = bytes([w[0] & 0x3F]) + w[1:]
            assert blob[32 * i: 32 * i + 32] == w


def test_parallel_stage_copy_is_a_faithful_memcpy(lib):
    """The host-thread pool that stages pageable caller buffers (csrc/lwkzg.cu HostStager) moves bytes and nothing else:
    odd sizes, sizes around the slice boundaries, back-to-back calls."""
    import random

    lib.lwkzg_debug_stage_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.lwkzg_debug_stage_copy.restype = None
    rng = random.Random(7)
    for size in (0, 1, 4095, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 3 * (1 << 20) + 12345, 33554432 + 7, 8 * 4096 * 9 + 3):
        src = (ctypes.c_ubyte * max(size, 1))()
        block = bytes(rng.getrandbits(8) for _ in range(4099))
        data = (block * (size // len(block) + 1))[:size]
        ctypes.memmove(src, data, size)
        dst = (ctypes.c_ubyte * (max(size, 1) + 16))()
        ctypes.memset(dst, 0xAB, size + 16)
        lib.lwkzg_debug_stage_copy(dst, src, size)
        assert bytes(dst[:size]) == data
        assert bytes(dst[size: size + 16]) == b"\xab" * 16   # nothing written past the end
