#!/usr/bin/env python3
"""BASELINE config 5: variable-base G1 MSM size sweep 2^12 .. 2^22 (synthetic points/scalars on the device).
Prints points/s and the fraction of the measured integer roofline per size (SURVEY §8d work model)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambdaworks_kzg_b200 as lw

lw.set_option("window_bits", int(os.environ.get("WB", "13")))
s = lw.load_trusted_setup_file(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trusted_setup.txt"))
peak = max(lw.imad_peak(0), lw.imad_peak(1))
rows = []
for lg in range(12, int(os.environ.get("MAXLG", "22")) + 1):
    n = 1 << lg
    ms, out = lw.bench_var_msm(n, s, iters=3 if lg < 20 else 2, seed=1)
    work = min((255 // c + 1) * (10 * n + 14 * (1 << c)) for c in range(4, 24)) + 256 * 9  # Fp mul, SURVEY §8d
    rows.append({"log2_n": lg, "ms": ms, "points_per_s": n / (ms * 1e-3), "roofline_frac": work * 300 / (ms * 1e-3) / peak, "result": out.hex()[:16]})
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"imad_peak_mac32_per_s": peak}))
