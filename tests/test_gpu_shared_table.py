"""GPU tier: "share_table" (SURVEY §8 f2's table cache, as it makes sense on this hardware): one digit table per
(SRS, window, device) -- shared by reference count inside a process and through a CUDA IPC handle between processes
-- gives byte-identical results and is attached instead of rebuilt (the loaders it short-cuts:
/root/reference/src/srs.rs:99-128, src/lib.rs:709-802)."""
import os
import subprocess
import sys
import time

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETUP = os.path.join(ROOT, "tests", "golden", "trusted_setup.txt")

CHILD = r"""
import sys, time
sys.path.insert(0, %r)
import lambdaworks_kzg_b200 as lw
lw.set_option("window_bits", 12)
lw.set_option("share_table", 1)
t0 = time.perf_counter()
s = lw.load_trusted_setup_file(%r)
dt = time.perf_counter() - t0
print("CHILD", lw.table_share(s), lw.window_bits(s), "%%.3f" %% dt, lw.blob_to_kzg_commitment(lw.synth_blob_host(7), s).hex())
s.free()
"""


def test_table_shared_in_process_and_across_processes():
    import lambdaworks_kzg_b200 as lw

    lw.load_library()
    lw.set_option("window_bits", 12)
    lw.set_option("share_table", 1)
    try:
        t0 = time.perf_counter()
        a = lw.load_trusted_setup_file(SETUP)
        t_build = time.perf_counter() - t0
        assert lw.table_share(a) == 1 and lw.window_bits(a) == 12
        t0 = time.perf_counter()
        b = lw.load_trusted_setup_file(SETUP)        # same process: reference-counted
        t_again = time.perf_counter() - t0
        assert lw.table_share(b) == 3 and lw.window_bits(b) == 12
        blob = lw.synth_blob_host(7)
        want = lw.blob_to_kzg_commitment(blob, a)
        assert lw.blob_to_kzg_commitment(blob, b) == want
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, SETUP)], capture_output=True, text=True, timeout=300)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("CHILD")]
        assert line, out.stdout + out.stderr
        _, share, wb, dt, com = line[0].split()
        assert int(share) == 2 and int(wb) == 12 and com == want.hex()
        print("build %.2f s, second settings %.2f s, second process load %s s" % (t_build, t_again, dt))
        a.free()                                      # b still holds the table
        assert lw.blob_to_kzg_commitment(blob, b) == want
        b.free()
        c = lw.load_trusted_setup_file(SETUP)        # last reference gone: rebuilt and published again
        assert lw.table_share(c) == 1 and lw.blob_to_kzg_commitment(blob, c) == want
        c.free()
    finally:
        lw.set_option("share_table", 0)
        lw.set_option("window_bits", 13)
    # a private table for comparison
    lw.set_option("window_bits", 12)
    try:
        d = lw.load_trusted_setup_file(SETUP)
        assert lw.table_share(d) == 0 and lw.blob_to_kzg_commitment(lw.synth_blob_host(7), d) == want
        d.free()
    finally:
        lw.set_option("window_bits", 13)
