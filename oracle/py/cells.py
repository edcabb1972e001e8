"""Python big-int oracle for the PeerDAS / EIP-7594 cell operations.

TEST INFRASTRUCTURE ONLY (see oracle/py/bls.py header).

PARITY UNPINNED: /root/reference implements none of this -- it carries the 65 G2
points such a path needs but only ever reads two (src/srs.rs:274, constants at
src/lib.rs:60-92) and holds no cell vectors.  This file restates consensus-specs
`specs/fulu/polynomial-commitments-sampling.md` (not part of /root/reference):
`compute_cells_and_kzg_proofs`, `verify_cell_kzg_proof_batch`,
`recover_cells_and_kzg_proofs` and their helpers, on top of the DenebMode / LeMode /
RefMode field encodings of oracle/py/kzg.py.  What pins it instead:

* proofs are computed WITHOUT FK20, from the toxic waste tau of tests/trusted_setup.txt:
  pi_k = [(p(tau) - I_k(tau)) / (tau^64 - h_k^64)] G, and (slow path, `generic=True`) as
  an explicit MSM of the quotient's coefficients over the monomial SRS;
* cells 0..63 of an extended blob are the blob itself; every cell is p evaluated on a
  coset, checked against Horner evaluation of the coefficient form;
* verification is the spec's universal equation with the pairing replaced by its
  toxic-waste form, RL == [tau^64] LL, and proofs made by this file must verify, proofs
  of a different cell / commitment must not.

Semantic modes: the cell API follows the mode of the settings it is called with.  In the
Lagrange modes (1 = little-endian c-kzg era, 2 = Deneb / mainnet big-endian) a blob is the
evaluation form over the bit-reversed 4096th roots of unity, as in the spec.  In
MODE_REFERENCE (0) a blob is what the reference says it is -- 4096 monomial coefficients,
big-endian, reduced mod r (src/utils.rs:27-41) -- so cells 0..63 are NOT the blob there.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Tuple

from . import bls
from .bls import R
from .kzg import (BYTES_PER_BLOB, C_KZG_BADARGS, C_KZG_ERROR, FIELD_ELEMENTS_PER_BLOB, PRIMITIVE_ROOT, KzgError, Setup, _bitrev, batch_inv)

FIELD_ELEMENTS_PER_EXT_BLOB = 2 * FIELD_ELEMENTS_PER_BLOB
FIELD_ELEMENTS_PER_CELL = 64
BYTES_PER_CELL = 32 * FIELD_ELEMENTS_PER_CELL
CELLS_PER_EXT_BLOB = FIELD_ELEMENTS_PER_EXT_BLOB // FIELD_ELEMENTS_PER_CELL  # 128
RANDOM_CHALLENGE_KZG_CELL_BATCH_DOMAIN = b"RCKZGCBATCH__V1_"


def root_of_unity(order: int) -> int:
    return pow(PRIMITIVE_ROOT, (R - 1) // order, R)


def fft(vals: Sequence[int], w: int) -> List[int]:
    """out[k] = sum_n vals[n] w^(k n); len(vals) a power of two, w a primitive len-th root."""
    n = len(vals)
    if n == 1:
        return list(vals)
    even = fft(vals[0::2], w * w % R)
    odd = fft(vals[1::2], w * w % R)
    out = [0] * n
    t = 1
    for k in range(n // 2):
        x = t * odd[k] % R
        out[k] = (even[k] + x) % R
        out[k + n // 2] = (even[k] - x) % R
        t = t * w % R
    return out


def ifft(vals: Sequence[int], w: int) -> List[int]:
    n = len(vals)
    ninv = bls.fr_inv(n)
    return [v * ninv % R for v in fft(vals, bls.fr_inv(w))]


def brp(seq: Sequence) -> List:
    bits = len(seq).bit_length() - 1
    return [seq[_bitrev(i, bits)] for i in range(len(seq))]


def coset_shift_for_cell(cell_index: int) -> int:
    """h_k: first element of coset_for_cell(k) = brp(roots of unity of order 8192)[64 k]."""
    return pow(root_of_unity(FIELD_ELEMENTS_PER_EXT_BLOB), _bitrev(FIELD_ELEMENTS_PER_CELL * cell_index, 13), R)


def coset_for_cell(cell_index: int) -> List[int]:
    w = root_of_unity(FIELD_ELEMENTS_PER_EXT_BLOB)
    return [pow(w, _bitrev(FIELD_ELEMENTS_PER_CELL * cell_index + j, 13), R) for j in range(FIELD_ELEMENTS_PER_CELL)]


def horner(coeffs: Sequence[int], z: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R
    return acc


class CellOracle:
    """mode: 0 = MODE_REFERENCE, 1 = MODE_CKZG_LE, 2 = MODE_DENEB (see module docstring)."""

    def __init__(self, setup: Setup, mode: int = 2, generic: bool = False):
        assert setup.tau is not None, "the cell oracle needs the toxic waste"
        self.s = setup
        self.mode = mode
        self.generic = generic
        self.bad = C_KZG_ERROR if mode == 0 else C_KZG_BADARGS

    # ---- field / point encodings of the mode
    def fr_from_bytes(self, b: bytes) -> int:
        if self.mode == 0:
            return int.from_bytes(b, "big") % R  # utils.rs:27-41 + App. A.1: reduced, never rejected
        v = int.from_bytes(b, "little" if self.mode == 1 else "big")
        if v >= R:
            raise KzgError(C_KZG_BADARGS, "non-canonical field element")
        return v

    def fr_to_bytes(self, v: int) -> bytes:
        return (v % R).to_bytes(32, "little" if self.mode == 1 else "big")

    def _decompress(self, b: bytes):
        try:
            return bls.g1_decompress(b, strict=self.mode != 0)
        except bls.PointError as e:
            raise KzgError(self.bad, str(e))

    # ---- polynomial forms
    def blob_to_coeffs(self, blob: bytes) -> List[int]:
        if len(blob) != BYTES_PER_BLOB:
            raise KzgError(C_KZG_BADARGS, "blob length")
        words = [self.fr_from_bytes(blob[i : i + 32]) for i in range(0, BYTES_PER_BLOB, 32)]
        if self.mode == 0:
            return words
        # polynomial_eval_to_coeff: ifft(bit_reversal_permutation(evaluations))
        return ifft(brp(words), root_of_unity(FIELD_ELEMENTS_PER_BLOB))

    def cells_from_coeffs(self, coeffs: Sequence[int]) -> List[List[int]]:
        ext = fft(list(coeffs) + [0] * FIELD_ELEMENTS_PER_BLOB, root_of_unity(FIELD_ELEMENTS_PER_EXT_BLOB))
        ext = brp(ext)
        return [ext[i * FIELD_ELEMENTS_PER_CELL : (i + 1) * FIELD_ELEMENTS_PER_CELL] for i in range(CELLS_PER_EXT_BLOB)]

    def cell_to_bytes(self, evals: Sequence[int]) -> bytes:
        return b"".join(self.fr_to_bytes(v) for v in evals)

    def cell_to_evals(self, cell: bytes) -> List[int]:
        if len(cell) != BYTES_PER_CELL:
            raise KzgError(C_KZG_BADARGS, "cell length")
        if self.mode == 0:  # cells are library OUTPUTS in canonical form: a non-canonical one is an error even here
            out = []
            for i in range(0, BYTES_PER_CELL, 32):
                v = int.from_bytes(cell[i : i + 32], "big")
                if v >= R:
                    raise KzgError(C_KZG_ERROR, "non-canonical field element")
                out.append(v)
            return out
        return [self.fr_from_bytes(cell[i : i + 32]) for i in range(0, BYTES_PER_CELL, 32)]

    # ---- proofs
    @staticmethod
    def interpolation_coeffs_from_poly(coeffs: Sequence[int], cell_index: int) -> List[int]:
        """I_k = p mod (X^64 - h_k^64): a_m = sum_u f_(64 u + m) (h_k^64)^u."""
        c = pow(coset_shift_for_cell(cell_index), FIELD_ELEMENTS_PER_CELL, R)
        out = []
        for m in range(FIELD_ELEMENTS_PER_CELL):
            acc = 0
            for u in reversed(range(len(coeffs) // FIELD_ELEMENTS_PER_CELL)):
                acc = (acc * c + coeffs[FIELD_ELEMENTS_PER_CELL * u + m]) % R
            out.append(acc)
        return out

    def proof_for_cell(self, coeffs: Sequence[int], cell_index: int):
        """compute_kzg_proof_multi_impl: [q(tau)]G, q = (p - I_k) / (X^64 - h_k^64)."""
        n = FIELD_ELEMENTS_PER_CELL
        c = pow(coset_shift_for_cell(cell_index), n, R)
        if self.generic:
            # explicit long division by X^64 - c, then an MSM over the monomial SRS
            rem = list(coeffs)
            q = [0] * (len(coeffs) - n)
            for i in range(len(coeffs) - 1, n - 1, -1):
                q[i - n] = rem[i]
                rem[i - n] = (rem[i - n] + c * rem[i]) % R
                rem[i] = 0
            return bls.g1_msm(self.s.g1[: len(q)], q)
        tau = self.s.tau
        interp = self.interpolation_coeffs_from_poly(coeffs, cell_index)
        num = (horner(coeffs, tau) - horner(interp, tau)) % R
        den = (pow(tau, n, R) - c) % R
        return bls.g1_mul(bls.G1, num * bls.fr_inv(den) % R)

    def compute_cells_and_kzg_proofs(self, blob: bytes, want_proofs: bool = True, cell_subset: Optional[Sequence[int]] = None) -> Tuple[List[bytes], List[bytes]]:
        coeffs = self.blob_to_coeffs(blob)
        cells = [self.cell_to_bytes(c) for c in self.cells_from_coeffs(coeffs)]
        proofs: List[bytes] = []
        if want_proofs:
            idx = range(CELLS_PER_EXT_BLOB) if cell_subset is None else cell_subset
            proofs = [bls.g1_compress(self.proof_for_cell(coeffs, k)) for k in idx]
        return cells, proofs

    # ---- verification (verify_cell_kzg_proof_batch / _impl of the spec)
    def batch_challenge(self, commitments: Sequence[bytes], commitment_indices, cell_indices, cells: Sequence[bytes], proofs: Sequence[bytes]) -> int:
        e = "little" if self.mode == 1 else "big"
        msg = RANDOM_CHALLENGE_KZG_CELL_BATCH_DOMAIN
        msg += FIELD_ELEMENTS_PER_BLOB.to_bytes(8, e) + FIELD_ELEMENTS_PER_CELL.to_bytes(8, e)
        msg += len(commitments).to_bytes(8, e) + len(cell_indices).to_bytes(8, e)
        msg += b"".join(commitments)
        for k in range(len(cell_indices)):
            msg += int(commitment_indices[k]).to_bytes(8, e) + int(cell_indices[k]).to_bytes(8, e) + cells[k] + proofs[k]
        return int.from_bytes(hashlib.sha256(msg).digest(), e) % R

    def verify_cell_kzg_proof_batch(self, commitments_bytes: Sequence[bytes], cell_indices: Sequence[int], cells: Sequence[bytes], proofs_bytes: Sequence[bytes]) -> bool:
        if not (len(commitments_bytes) == len(cell_indices) == len(cells) == len(proofs_bytes)):
            raise KzgError(C_KZG_BADARGS, "length mismatch")
        for c in commitments_bytes:
            if len(c) != 48:
                raise KzgError(C_KZG_BADARGS, "length")
        for ci in cell_indices:
            if not 0 <= ci < CELLS_PER_EXT_BLOB:
                raise KzgError(C_KZG_BADARGS, "cell index")
        for p in proofs_bytes:
            if len(p) != 48:
                raise KzgError(C_KZG_BADARGS, "length")
        if len(cells) == 0:
            return True
        # deduplicated commitments, in order of first appearance
        uniq: List[bytes] = []
        cidx = []
        for c in commitments_bytes:
            if c not in uniq:
                uniq.append(c)
            cidx.append(uniq.index(c))
        evals = [self.cell_to_evals(c) for c in cells]
        cpts = [self._decompress(c) for c in uniq]
        ppts = [self._decompress(p) for p in proofs_bytes]
        n = FIELD_ELEMENTS_PER_CELL
        r = self.batch_challenge(uniq, cidx, cell_indices, cells, proofs_bytes)
        rp = [pow(r, k, R) for k in range(len(cells))]
        ll = bls.g1_sum(bls.g1_mul(ppts[k], rp[k]) for k in range(len(cells)))
        weights = [0] * len(uniq)
        for k, i in enumerate(cidx):
            weights[i] = (weights[i] + rp[k]) % R
        rlc = bls.g1_sum(bls.g1_mul(cpts[i], weights[i]) for i in range(len(uniq)))
        w64 = root_of_unity(n)
        summed = [0] * n
        for k in range(len(cells)):
            h = coset_shift_for_cell(cell_indices[k])
            # interpolate_polynomialcoeff over the coset h * <w64>, evaluations given in bit-reversed order
            a = ifft(brp(evals[k]), w64)
            hinv = bls.fr_inv(h)
            t = 1
            for m in range(n):
                summed[m] = (summed[m] + rp[k] * a[m] % R * t) % R
                t = t * hinv % R
        rli = bls.g1_msm(self.s.g1[:n], summed) if self.generic else bls.g1_mul(bls.G1, horner(summed, self.s.tau))
        rlp = bls.g1_sum(bls.g1_mul(ppts[k], rp[k] * pow(coset_shift_for_cell(cell_indices[k]), n, R) % R) for k in range(len(cells)))
        rl = bls.g1_add(bls.g1_add(rlc, bls.g1_neg(rli)), rlp)
        if self.generic:
            from . import pairing

            return pairing.pairing_product_is_one([(ll, self.s.g2[n]), (bls.g1_neg(rl), self.s.g2[0])])
        return rl == bls.g1_mul(ll, pow(self.s.tau, n, R))

    # ---- recovery (recover_cells_and_kzg_proofs of the spec)
    def recover_coeffs(self, cell_indices: Sequence[int], cells_evals: Sequence[Sequence[int]]) -> List[int]:
        """recover_polynomialcoeff: vanishing polynomial of the missing cells, (E Z)(x) on the whole domain, division on
        a shifted coset."""
        n_ext, n_cell = FIELD_ELEMENTS_PER_EXT_BLOB, FIELD_ELEMENTS_PER_CELL
        w_ext = root_of_unity(n_ext)
        w_cells = root_of_unity(CELLS_PER_EXT_BLOB)
        missing = [i for i in range(CELLS_PER_EXT_BLOB) if i not in set(cell_indices)]
        # short vanishing polynomial over the 128th roots w_cells^brp(i), then X -> X^64
        short = [1]
        for i in missing:
            root = pow(w_cells, _bitrev(i, 7), R)
            nxt = [0] * (len(short) + 1)
            for d, cf in enumerate(short):
                nxt[d + 1] = (nxt[d + 1] + cf) % R
                nxt[d] = (nxt[d] - cf * root) % R
            short = nxt
        zero_poly = [0] * n_ext
        for d, cf in enumerate(short):
            zero_poly[d * n_cell] = cf
        ext_brp = [0] * n_ext
        for ci, ev in zip(cell_indices, cells_evals):
            ext_brp[ci * n_cell : (ci + 1) * n_cell] = list(ev)
        ext = brp(ext_brp)
        zero_eval = fft(zero_poly, w_ext)
        ez = [a * b % R for a, b in zip(ext, zero_eval)]
        ez_coeff = ifft(ez, w_ext)
        shift = PRIMITIVE_ROOT
        def coset_fft(cf):
            t, out = 1, []
            for v in cf:
                out.append(v * t % R)
                t = t * shift % R
            return fft(out, w_ext)
        num = coset_fft(ez_coeff)
        den = coset_fft(zero_poly)
        quo = [a * b % R for a, b in zip(num, batch_inv(den))]
        cf = ifft(quo, w_ext)
        sinv = bls.fr_inv(shift)
        t, out = 1, []
        for v in cf:
            out.append(v * t % R)
            t = t * sinv % R
        assert all(v == 0 for v in out[FIELD_ELEMENTS_PER_BLOB:]), "recovered polynomial has degree >= 4096"
        return out[:FIELD_ELEMENTS_PER_BLOB]

    def recover_cells_and_kzg_proofs(self, cell_indices: Sequence[int], cells: Sequence[bytes], want_proofs: bool = True, cell_subset: Optional[Sequence[int]] = None):
        if len(cell_indices) != len(cells):
            raise KzgError(C_KZG_BADARGS, "length mismatch")
        if not CELLS_PER_EXT_BLOB // 2 <= len(cell_indices) <= CELLS_PER_EXT_BLOB:
            raise KzgError(C_KZG_BADARGS, "need at least half of the cells")
        if len(set(cell_indices)) != len(cell_indices):
            raise KzgError(C_KZG_BADARGS, "duplicate cell index")
        for ci in cell_indices:
            if not 0 <= ci < CELLS_PER_EXT_BLOB:
                raise KzgError(C_KZG_BADARGS, "cell index")
        if list(cell_indices) != sorted(cell_indices):
            raise KzgError(C_KZG_BADARGS, "cell indices must ascend")  # c-kzg-4844 v2 requires ascending order
        evals = [self.cell_to_evals(c) for c in cells]
        coeffs = self.recover_coeffs(cell_indices, evals)
        out_cells = [self.cell_to_bytes(c) for c in self.cells_from_coeffs(coeffs)]
        proofs: List[bytes] = []
        if want_proofs:
            idx = range(CELLS_PER_EXT_BLOB) if cell_subset is None else cell_subset
            proofs = [bls.g1_compress(self.proof_for_cell(coeffs, k)) for k in idx]
        return out_cells, proofs
