"""Python big-int oracle for the EIP-4844 hot path of lambdaworks_kzg.

TEST INFRASTRUCTURE ONLY (see oracle/py/bls.py header).

Two semantic modes (SURVEY.md §0 finding 3/4, App. A and App. B):

* ``RefMode`` -- what the reference actually computes: big-endian scalars
  reduced mod r, blob words are monomial *coefficients*, the SRS is used as
  loaded, Horner evaluation, Ruffini division.  Follows
  /root/reference/src/lib.rs:253-692, src/utils.rs:27-206,
  src/compression.rs:22-139, src/srs.rs:25-128.
* ``LeMode`` -- the little-endian-era c-kzg-4844 semantics that the YAML
  vectors under /root/reference/tests/*/small encode (never loaded by the
  reference itself).  Used to pin this oracle's curve arithmetic, SHA layout
  and codecs against 208 third-party known answers.

Both modes offer a "toxic waste" shortcut when the setup's secret tau is known
(tests/trusted_setup.txt has tau = 1337): commit(c) = [sum c_i tau^i]G and
verify(C,z,y,pi) <=> C - yG == [tau - z]pi, which is independent of any MSM or
pairing code.  ``generic=True`` forces the MSM path.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import bls
from .bls import P, R, PointError

FIELD_ELEMENTS_PER_BLOB = 4096
BYTES_PER_BLOB = 4096 * 32
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"  # lib.rs:60
RANDOM_CHALLENGE_KZG_BATCH_DOMAIN = b"RCKZGBATCH___V1_"  # lib.rs:62

C_KZG_OK, C_KZG_BADARGS, C_KZG_ERROR, C_KZG_MALLOC = 0, 1, 2, 3


class KzgError(Exception):
    def __init__(self, code=C_KZG_ERROR, msg=""):
        super().__init__(msg)
        self.code = code


@dataclass
class Setup:
    g1: list  # affine tuples / None, monomial order as in the file
    g2: list
    tau: Optional[int] = None
    _lagrange_brp: Optional[list] = field(default=None, repr=False)


def parse_setup_text(text: str, *, check_subgroup: bool = False) -> Setup:
    """src/srs.rs:25-82: line 1 = n1, line 2 = n2, then n1 G1 hex lines and n2
    G2 hex lines.  check_subgroup=False skips the 4096 [r]P checks (slow in
    Python); the C oracle performs them."""
    lines = text.splitlines()
    n1, n2 = int(lines[0]), int(lines[1])
    g1 = []
    for ln in lines[2 : 2 + n1]:
        raw = bytes.fromhex(ln.strip())
        if check_subgroup:
            g1.append(bls.g1_decompress(raw))
        else:
            g1.append(_g1_decompress_nocheck(raw))
    g2 = [bls.g2_decompress(bytes.fromhex(ln.strip())) for ln in lines[2 + n1 : 2 + n1 + n2]]
    s = Setup(g1=g1, g2=g2)
    s.tau = _detect_tau(s)
    return s


def _g1_decompress_nocheck(raw: bytes):
    b0 = raw[0]
    if not b0 & 0x80:
        raise PointError("not compressed")
    if b0 & 0x40:
        return None
    x = int.from_bytes(bytes([b0 & 0x1F]) + raw[1:], "big") % P
    y = bls.fp_sqrt(x * x * x + 4)
    if y is None:
        raise PointError("not on curve")
    lo, hi = (y, P - y) if y < P - y else (P - y, y)
    return (x, hi if b0 & 0x20 else lo)


def _detect_tau(s: Setup) -> Optional[int]:
    """Return 1337 iff the setup is the consensus-specs testing setup."""
    tau = 1337
    if len(s.g1) < 2 or s.g1[0] != bls.G1:
        return None
    if s.g1[1] != bls.g1_mul(bls.G1, tau):
        return None
    last = len(s.g1) - 1
    if s.g1[last] != bls.g1_mul(bls.G1, pow(tau, last, R)):
        return None
    return tau


# ===================================================================== shared


def sha256(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


def horner(coeffs: Sequence[int], z: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R
    return acc


def ruffini(coeffs: Sequence[int], z: int) -> List[int]:
    """(p - p(z)) / (X - z); len = len(coeffs) - 1."""
    n = len(coeffs)
    if n <= 1:
        return []
    q = [0] * (n - 1)
    acc = 0
    for k in range(n - 1, 0, -1):
        acc = (acc * z + coeffs[k]) % R
        q[k - 1] = acc
    return q


def _trim(coeffs: List[int]) -> List[int]:
    n = len(coeffs)
    while n and coeffs[n - 1] == 0:
        n -= 1
    return coeffs[:n]


# =================================================================== RefMode


class RefMode:
    """MODE_REFERENCE."""

    def __init__(self, setup: Setup, generic: bool = False):
        self.s = setup
        self.generic = generic or setup.tau is None

    # -- parsing (utils.rs:27-41; App. A.1/A.2)
    @staticmethod
    def blob_to_coeffs(blob: bytes) -> List[int]:
        assert len(blob) == BYTES_PER_BLOB
        return _trim([int.from_bytes(blob[i : i + 32], "big") % R for i in range(0, BYTES_PER_BLOB, 32)])

    @staticmethod
    def fr_from_bytes(b: bytes) -> int:
        return int.from_bytes(b, "big") % R

    @staticmethod
    def fr_to_bytes(v: int) -> bytes:
        return (v % R).to_bytes(32, "big")

    def _check_srs(self):
        # srs.rs:155-172 + 258-280: every g1 value must be on the curve
        # (infinity (0,0) fails from_affine) -- otherwise every call errors.
        if any(p is None for p in self.s.g1):
            raise KzgError(C_KZG_ERROR, "SRS re-hydration failed")

    # -- commit (lib.rs:266-270; App. A.3)
    def commit_coeffs(self, coeffs: Sequence[int]):
        self._check_srs()
        if not coeffs:
            return None
        if not self.generic:
            tau = self.s.tau
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * tau + c) % R
            return bls.g1_mul(bls.G1, acc)
        return bls.g1_msm(self.s.g1[: len(coeffs)], list(coeffs))

    def blob_to_kzg_commitment(self, blob: bytes) -> bytes:
        return bls.g1_compress(self.commit_coeffs(self.blob_to_coeffs(blob)))

    # -- proofs (lib.rs:300-404; App. A.4/A.5)
    def _open(self, coeffs, z):
        y = horner(coeffs, z)
        q = _trim(ruffini(coeffs, z))
        return self.commit_coeffs(q), y

    def compute_kzg_proof(self, blob: bytes, z_bytes: bytes) -> Tuple[bytes, bytes]:
        coeffs = self.blob_to_coeffs(blob)
        z = self.fr_from_bytes(z_bytes)
        proof, y = self._open(coeffs, z)
        return bls.g1_compress(proof), self.fr_to_bytes(y)

    def compute_challenge(self, blob: bytes, commitment_pt) -> int:
        msg = (
            FIAT_SHAMIR_PROTOCOL_DOMAIN
            + (FIELD_ELEMENTS_PER_BLOB).to_bytes(8, "little")
            + (0).to_bytes(8, "little")
            + blob
            + bls.g1_compress(commitment_pt)
        )
        return int.from_bytes(sha256(msg), "big") % R

    def _decompress(self, b: bytes):
        try:
            return bls.g1_decompress(b)
        except PointError as e:
            raise KzgError(C_KZG_ERROR, str(e))

    def compute_blob_kzg_proof(self, blob: bytes, commitment_bytes: bytes) -> bytes:
        c = self._decompress(commitment_bytes)
        coeffs = self.blob_to_coeffs(blob)
        z = self.compute_challenge(blob, c)
        proof, _ = self._open(coeffs, z)
        return bls.g1_compress(proof)

    # -- verification (lib.rs:407-692; App. A.6/A.7)
    def _pairing_check(self, a1, b1) -> bool:
        """e(a1, g2[0]) * e(-b1, g2[1]) == 1  with g2[1] = [tau]g2[0]."""
        if self.generic:
            from . import pairing

            return pairing.pairing_product_is_one([(a1, self.s.g2[0]), (bls.g1_neg(b1), self.s.g2[1])])
        return a1 == bls.g1_mul(b1, self.s.tau)

    def _verify(self, c, z: int, y: int, proof) -> bool:
        self._check_srs()
        # e(C - y g1[0], g2[0]) * e(-pi, g2[1] - z g2[0]) == 1
        #   <=> e(C - y g1[0] + z pi, g2[0]) == e(pi, g2[1])
        lhs = bls.g1_add(bls.g1_add(c, bls.g1_neg(bls.g1_mul(self.s.g1[0], y))), bls.g1_mul(proof, z))
        return self._pairing_check(lhs, proof)

    def verify_kzg_proof(self, commitment_bytes, z_bytes, y_bytes, proof_bytes) -> bool:
        c = self._decompress(commitment_bytes)
        z = self.fr_from_bytes(z_bytes)
        y = self.fr_from_bytes(y_bytes)
        pi = self._decompress(proof_bytes)
        return self._verify(c, z, y, pi)

    def verify_blob_kzg_proof(self, blob, commitment_bytes, proof_bytes) -> bool:
        c = self._decompress(commitment_bytes)
        pi = self._decompress(proof_bytes)
        coeffs = self.blob_to_coeffs(blob)
        z = self.compute_challenge(blob, c)
        y = horner(coeffs, z)
        return self._verify(c, z, y, pi)

    def batch_challenge(self, cs, zs, ys, pis) -> int:
        n = len(cs)
        msg = RANDOM_CHALLENGE_KZG_BATCH_DOMAIN + (FIELD_ELEMENTS_PER_BLOB).to_bytes(8, "little") + n.to_bytes(8, "little")
        for c, z, y, pi in zip(cs, zs, ys, pis):
            msg += bls.g1_compress(c) + self.fr_to_bytes(z) + self.fr_to_bytes(y) + bls.g1_compress(pi)
        return int.from_bytes(sha256(msg), "big") % R

    def verify_blob_kzg_proof_batch(self, blobs: Sequence[bytes], commitments: Sequence[bytes], proofs: Sequence[bytes]) -> bool:
        n = len(blobs)
        if n == 0:
            return False  # lib.rs:538-543
        if n == 1:
            return self.verify_blob_kzg_proof(blobs[0], commitments[0], proofs[0])
        cs, zs, ys, pis = [], [], [], []
        for i in range(n):  # lib.rs:562-596 (order of checks)
            c = self._decompress(commitments[i])
            coeffs = self.blob_to_coeffs(blobs[i])
            z = self.compute_challenge(blobs[i], c)
            ys.append(horner(coeffs, z))
            zs.append(z)
            cs.append(c)
            pis.append(self._decompress(proofs[i]))
        return self.verify_kzg_proof_batch(cs, zs, ys, pis)

    def verify_kzg_proof_batch(self, cs, zs, ys, pis) -> bool:
        self._check_srs()
        n = len(cs)
        r = self.batch_challenge(cs, zs, ys, pis)
        rp = [pow(r, i, R) for i in range(n)]
        G = bls.G1  # lib.rs:661 uses the curve generator, not g1_values[0]
        c_minus_y = [bls.g1_add(cs[i], bls.g1_neg(bls.g1_mul(G, ys[i]))) for i in range(n)]
        proof_lincomb = bls.g1_sum(bls.g1_mul(pis[i], rp[i]) for i in range(n))
        proof_z_lincomb = bls.g1_sum(bls.g1_mul(pis[i], rp[i] * zs[i] % R) for i in range(n))
        c_minus_y_lincomb = bls.g1_sum(bls.g1_mul(c_minus_y[i], rp[i]) for i in range(n))
        rhs = bls.g1_add(c_minus_y_lincomb, proof_z_lincomb)
        return self._pairing_check(rhs, proof_lincomb)


# ==================================================================== LeMode

PRIMITIVE_ROOT = 7


def _bitrev(i: int, bits: int) -> int:
    return int(format(i, "0%db" % bits)[::-1], 2)


_DOMAIN_CACHE = {}


def brp_domain(n: int = FIELD_ELEMENTS_PER_BLOB) -> List[int]:
    if n not in _DOMAIN_CACHE:
        w = pow(PRIMITIVE_ROOT, (R - 1) // n, R)
        bits = n.bit_length() - 1
        pw = [1] * n
        for i in range(1, n):
            pw[i] = pw[i - 1] * w % R
        _DOMAIN_CACHE[n] = [pw[_bitrev(i, bits)] for i in range(n)]
    return _DOMAIN_CACHE[n]


def batch_inv(vals: Sequence[int]) -> List[int]:
    n = len(vals)
    pre = [1] * (n + 1)
    for i, v in enumerate(vals):
        pre[i + 1] = pre[i] * v % R
    inv = bls.fr_inv(pre[n])
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = pre[i] * inv % R
        inv = inv * vals[i] % R
    return out


class LeMode:
    """MODE_CKZG_LE: little-endian canonical scalars, blob =
    evaluations over the bit-reversed 4096th roots of unity."""

    def __init__(self, setup: Setup):
        assert setup.tau is not None, "LeMode oracle needs the toxic waste"
        self.s = setup
        self.dom = brp_domain()

    @staticmethod
    def fr_from_bytes(b: bytes) -> int:
        v = int.from_bytes(b, "little")
        if v >= R:
            raise KzgError(C_KZG_BADARGS, "non-canonical field element")
        return v

    @staticmethod
    def fr_to_bytes(v: int) -> bytes:
        return (v % R).to_bytes(32, "little")

    def blob_to_evals(self, blob: bytes) -> List[int]:
        if len(blob) != BYTES_PER_BLOB:
            raise KzgError(C_KZG_BADARGS, "blob length")
        return [self.fr_from_bytes(blob[i : i + 32]) for i in range(0, BYTES_PER_BLOB, 32)]

    def eval_at(self, evals: Sequence[int], z: int) -> int:
        n = len(evals)
        for i, w in enumerate(self.dom):
            if w == z:
                return evals[i]
        inv = batch_inv([(z - w) % R for w in self.dom])
        acc = 0
        for e, w, iv in zip(evals, self.dom, inv):
            acc = (acc + e * w % R * iv) % R
        return acc * (pow(z, n, R) - 1) % R * bls.fr_inv(n) % R

    def _decompress(self, b: bytes):
        if len(b) != 48:
            raise KzgError(C_KZG_BADARGS, "length")
        try:
            return bls.g1_decompress(b, strict=True)
        except PointError as e:
            raise KzgError(C_KZG_BADARGS, str(e))

    def blob_to_kzg_commitment(self, blob: bytes) -> bytes:
        evals = self.blob_to_evals(blob)
        return bls.g1_compress(bls.g1_mul(bls.G1, self.eval_at(evals, self.s.tau)))

    def _proof(self, evals, z) -> Tuple[bytes, int]:
        y = self.eval_at(evals, z)
        tau = self.s.tau
        q_tau = (self.eval_at(evals, tau) - y) * bls.fr_inv((tau - z) % R) % R
        return bls.g1_compress(bls.g1_mul(bls.G1, q_tau)), y

    def compute_kzg_proof(self, blob: bytes, z_bytes: bytes):
        evals = self.blob_to_evals(blob)
        if len(z_bytes) != 32:
            raise KzgError(C_KZG_BADARGS, "length")
        z = self.fr_from_bytes(z_bytes)
        proof, y = self._proof(evals, z)
        return proof, self.fr_to_bytes(y)

    def compute_challenge(self, blob: bytes, commitment_bytes: bytes) -> int:
        msg = (
            FIAT_SHAMIR_PROTOCOL_DOMAIN
            + (FIELD_ELEMENTS_PER_BLOB).to_bytes(8, "little")
            + (0).to_bytes(8, "little")
            + blob
            + commitment_bytes
        )
        return int.from_bytes(sha256(msg), "little") % R

    def compute_blob_kzg_proof(self, blob: bytes, commitment_bytes: bytes) -> bytes:
        evals = self.blob_to_evals(blob)
        self._decompress(commitment_bytes)
        z = self.compute_challenge(blob, commitment_bytes)
        return self._proof(evals, z)[0]

    def _verify(self, c, z, y, pi) -> bool:
        tau = self.s.tau
        lhs = bls.g1_add(c, bls.g1_neg(bls.g1_mul(bls.G1, y)))
        return lhs == bls.g1_mul(pi, (tau - z) % R)

    def verify_kzg_proof(self, commitment_bytes, z_bytes, y_bytes, proof_bytes) -> bool:
        c = self._decompress(commitment_bytes)
        if len(z_bytes) != 32 or len(y_bytes) != 32:
            raise KzgError(C_KZG_BADARGS, "length")
        z = self.fr_from_bytes(z_bytes)
        y = self.fr_from_bytes(y_bytes)
        pi = self._decompress(proof_bytes)
        return self._verify(c, z, y, pi)

    def verify_blob_kzg_proof(self, blob, commitment_bytes, proof_bytes) -> bool:
        evals = self.blob_to_evals(blob)
        c = self._decompress(commitment_bytes)
        pi = self._decompress(proof_bytes)
        z = self.compute_challenge(blob, commitment_bytes)
        y = self.eval_at(evals, z)
        return self._verify(c, z, y, pi)

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs) -> bool:
        if not (len(blobs) == len(commitments) == len(proofs)):
            raise KzgError(C_KZG_BADARGS, "length mismatch")
        ok = True
        for b, c, p in zip(blobs, commitments, proofs):
            # every element is validated (errors win over a False result)
            ok = self.verify_blob_kzg_proof(b, c, p) and ok
        return ok


# ================================================================= DenebMode


class DenebMode(LeMode):
    """MODE_DENEB: the final (mainnet) EIP-4844 wire format -- the combination the
    reference stopped short of: it parses big-endian scalars
    (/root/reference/src/utils.rs:27-41) but left the Lagrange conversion of the
    SRS as a TODO (/root/reference/src/lib.rs:760-770, src/srs.rs:117-124).

    Restates consensus-specs `specs/deneb/polynomial-commitments.md` (not part of
    /root/reference; KZG_ENDIANNESS = 'big'): `bytes_to_bls_field` (canonical
    big-endian, >= r rejected), `blob_to_polynomial` (evaluation form over the
    bit-reversed roots of unity), `compute_challenge` (domain || 16-byte big-endian
    degree || blob || commitment, digest read big-endian mod r),
    `evaluate_polynomial_in_evaluation_form`, `compute_kzg_proof_impl`
    (`compute_quotient_eval_within_domain` when z is a root of unity),
    `verify_kzg_proof_impl`, `verify_kzg_proof_batch` / `compute_powers`.
    Everything that does not touch a hash is the LeMode computation on byte-reversed
    field elements, which the 208 YAML vectors pin (tests/test_oracle.py)."""

    @staticmethod
    def fr_from_bytes(b: bytes) -> int:
        v = int.from_bytes(b, "big")
        if v >= R:
            raise KzgError(C_KZG_BADARGS, "non-canonical field element")
        return v

    @staticmethod
    def fr_to_bytes(v: int) -> bytes:
        return (v % R).to_bytes(32, "big")

    def compute_challenge(self, blob: bytes, commitment_bytes: bytes) -> int:
        msg = FIAT_SHAMIR_PROTOCOL_DOMAIN + (FIELD_ELEMENTS_PER_BLOB).to_bytes(16, "big") + blob + commitment_bytes
        return int.from_bytes(sha256(msg), "big") % R

    def batch_challenge(self, commitments, zs, ys, proofs) -> int:
        n = len(commitments)
        msg = RANDOM_CHALLENGE_KZG_BATCH_DOMAIN + (FIELD_ELEMENTS_PER_BLOB).to_bytes(8, "big") + n.to_bytes(8, "big")
        for c, z, y, p in zip(commitments, zs, ys, proofs):
            msg += c + self.fr_to_bytes(z) + self.fr_to_bytes(y) + p
        return int.from_bytes(sha256(msg), "big") % R

    def verify_blob_kzg_proof_batch(self, blobs, commitments, proofs) -> bool:
        """verify_blob_kzg_proof_batch -> verify_kzg_proof_batch of the spec, with the
        random linear combination spelled out (toxic-waste form of the pairing check)."""
        if not (len(blobs) == len(commitments) == len(proofs)):
            raise KzgError(C_KZG_BADARGS, "length mismatch")
        cs, zs, ys, pis = [], [], [], []
        for b, c, p in zip(blobs, commitments, proofs):
            evals = self.blob_to_evals(b)
            cs.append(self._decompress(c))
            pis.append(self._decompress(p))
            z = self.compute_challenge(b, c)
            zs.append(z)
            ys.append(self.eval_at(evals, z))
        n = len(blobs)
        if n == 0:
            return True
        r = self.batch_challenge(commitments, zs, ys, proofs)
        rp = [pow(r, i, R) for i in range(n)]
        proof_lincomb = bls.g1_sum(bls.g1_mul(pis[i], rp[i]) for i in range(n))
        proof_z_lincomb = bls.g1_sum(bls.g1_mul(pis[i], rp[i] * zs[i] % R) for i in range(n))
        c_minus_y = [bls.g1_add(cs[i], bls.g1_neg(bls.g1_mul(bls.G1, ys[i]))) for i in range(n)]
        c_minus_y_lincomb = bls.g1_sum(bls.g1_mul(c_minus_y[i], rp[i]) for i in range(n))
        # e(proof_lincomb, -[tau]G2) * e(c_minus_y_lincomb + proof_z_lincomb, G2) == 1
        return bls.g1_add(c_minus_y_lincomb, proof_z_lincomb) == bls.g1_mul(proof_lincomb, self.s.tau)
