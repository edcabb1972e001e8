#!/usr/bin/env python3
"""Run under ncu (--metrics gpu__time_duration.sum) to get the per-kernel times of the variable-base MSM pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_kzg_b200 as lw

lw.set_option("window_bits", 8)
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
for lg in [int(x) for x in os.environ.get("LGS", "12,16").split(",")]:
    ms, out = lw.bench_var_msm(1 << lg, s, iters=1, seed=1)
    print(lg, ms, out.hex()[:16], flush=True)
