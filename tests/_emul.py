"""Build + load the host-emulation library of the device math headers."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cpp")
LIB = os.path.join(HERE, "host_emul", "libemul.so")
CSRC = os.path.join(os.path.dirname(HERE), "lambdaworks_kzg_b200", "csrc")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    EXP = os.path.join(os.path.dirname(HERE), "tools", "experiments")
    deps = [SRC] + [os.path.join(d, f) for d in (CSRC, EXP) for f in os.listdir(d) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        flags = ["-DLWKZG_HOST_EMUL"]
        if os.path.exists(os.path.join(CSRC, "pairing.cuh")):
            flags.append("-DLWKZG_EMUL_PAIRING")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def u32(v, n):
    return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def from_u32(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def aff_bytes(pt):
    if pt is None:
        return bytes(96)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def aff_from(b):
    b = bytes(b)
    if not any(b):
        return None
    return (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big"))
