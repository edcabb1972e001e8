"""Stage-by-stage check of the FK20 pipeline on the GPU against a model in the exponent (tau = 1337): prints the
first stage that deviates.  Run on the GPU box: python tools/cells_debug.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lambdaworks_kzg_b200 as lw  # noqa: E402
from oracle.py import bls, cells, kzg  # noqa: E402
from oracle.py.kzg import _bitrev  # noqa: E402
from tests.golden.make_cell_fixtures import make_blob  # noqa: E402

R, TAU = bls.R, 1337


def main():
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lw.set_option("mode", mode)
    lw.set_option("window_bits", 8)
    lw.set_option("cell_window_bits", 8)
    s = lw.load_trusted_setup_file(os.path.join(ROOT, "tests", "golden", "trusted_setup.txt"))
    setup = kzg.parse_setup_text(open(os.path.join(ROOT, "tests", "golden", "trusted_setup.txt")).read())
    o = cells.CellOracle(setup, mode)
    blob = make_blob(11, mode)
    f = o.blob_to_coeffs(blob)
    scalars, hhat, h, fk = lw.debug_cell_stages(blob, s)
    nu = pow(cells.root_of_unity(8192), 64, R)
    sp = [pow(TAU, k, R) for k in range(4096)]
    X, C = [], []
    for b in range(64):
        X.append(cells.fft([sp[64 * (62 - v) + b] if v <= 62 else 0 for v in range(128)], nu))
        c = [0] * 128
        c[0] = f[64 * 63 + b]
        for u in range(66, 128):
            c[u] = f[64 * (u - 65) + b]
        C.append(cells.fft(c, nu))
    inv128 = bls.fr_inv(128)
    bad = [(j, b) for j in range(128) for b in range(64) if scalars[j * 64 + b] != C[b][j] * inv128 % R]
    print("scalars: %d mismatches" % len(bad), bad[:4])
    badx = [(j, b) for j in range(128) for b in range(64) if fk[j * 64 + b] != bls.g1_mul(bls.G1, X[b][j])]
    print("fk20 points: %d mismatches" % len(badx), badx[:4])
    if badx:
        got = fk[0]
        print(" fk[0,0] =", got, " infinities:", sum(1 for q in fk if q is None))
        print(" fk[0,0] on curve:", got is not None and bls.g1_on_curve(got), " == input s_(64*62):", got == bls.g1_mul(bls.G1, sp[64 * 62]),
              " == G:", got == bls.G1)
        # which scalar is it?  try small combinations of the column b = 0
        col = [sp[64 * (62 - v)] for v in range(63)]
        for name, val in (("sum", sum(col) % R), ("alt", sum((-1) ** v * col[v] for v in range(63)) % R), ("first", col[0]), ("last", col[62])):
            print("  ", name, got == bls.g1_mul(bls.G1, val))
        good_j = sorted(set(j for j in range(128)) - set(j for j, b in badx if b == 0))
        print("  frequencies j with a correct point (b = 0):", good_j[:16])
    hh = [sum(C[b][j] * X[b][j] for b in range(64)) * inv128 % R for j in range(128)]
    badh = [j for j in range(128) if hhat[j] != bls.g1_compress(bls.g1_mul(bls.G1, hh[j]))]
    print("Hhat: %d mismatches" % len(badh), badh[:8])
    H = cells.fft(hh, bls.fr_inv(nu))
    badH = [p for p in range(128) if h[p] != bls.g1_compress(bls.g1_mul(bls.G1, H[_bitrev(p, 7)]) if p % 2 == 0 else None)]
    print("H (after inverse FFT): %d mismatches" % len(badH), badH[:8])
    cs, ps = lw.compute_cells_and_kzg_proofs(blob, s)
    proofs_nat = cells.fft(H[:64] + [0] * 64, nu)
    badp = [i for i in range(128) if ps[i] != bls.g1_compress(bls.g1_mul(bls.G1, proofs_nat[_bitrev(i, 7)]))]
    print("proofs: %d mismatches" % len(badp), badp[:8])
    s.free()


if __name__ == "__main__":
    main()
