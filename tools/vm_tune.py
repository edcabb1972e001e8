#!/usr/bin/env python3
"""Window size / lanes-per-bucket sweep of the variable-base MSM (LWKZG_VM_C, LWKZG_VM_S, LWKZG_VM_GLV are read per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_kzg_b200 as lw

lw.set_option("window_bits", 8)
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
for lg in [int(x) for x in os.environ.get("LGS", "12,13,14,15,16,17,18,19,20").split(",")]:
    n = 1 << lg
    ref = None
    rows = []
    for glv in (1, 0):
        os.environ["LWKZG_VM_GLV"] = str(glv)
        for c in range(6, 17):
            if not glv and c < lg - 5:
                continue
            if glv and c > lg:
                continue
            os.environ["LWKZG_VM_C"] = str(c)
            for S in ([0] if lg > 16 else [0, 2, 4, 8, 16]):
                if S:
                    os.environ["LWKZG_VM_S"] = str(S)
                else:
                    os.environ.pop("LWKZG_VM_S", None)
                ms, out = lw.bench_var_msm(n, s, iters=3 if lg < 18 else 2, seed=1)
                ref = ref or out
                assert out == ref, (lg, glv, c, S)
                rows.append((ms, glv, c, S))
    rows.sort()
    print("2^%d best:" % lg, ["%.3f ms glv=%d c=%d S=%d" % r for r in rows[:6]], flush=True)
