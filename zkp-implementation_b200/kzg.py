"""Host-side mirror of the reference's `kzg` crate surface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Rust API so the parity tests read like the
reference's own (kzg/src/commitment.rs:31-119):

=====================================  ===========================================
reference (kzg/src)                    here
=====================================  ===========================================
``Srs::new_from_secret``  srs.rs:48    :meth:`Srs.new_from_secret` (GPU fixed-base)
``Srs::g1_points``        srs.rs:78    :meth:`Srs.g1_points`
``KzgScheme::new``        scheme.rs:34 :class:`KzgScheme` (uploads the SRS once)
``commit``                scheme.rs:49 :meth:`KzgScheme.commit`
``commit_vector``         scheme.rs:63 :meth:`KzgScheme.commit_vector`
``commit_para``           scheme.rs:78 :meth:`KzgScheme.commit_para`
``open`` / ``open_vector``  :108/:132  :meth:`KzgScheme.open` / ``open_vector``
``KzgCommitment``   commitment.rs:5    :class:`KzgCommitment`
``KzgOpening``         opening.rs:12   :class:`KzgOpening`
=====================================  ===========================================

``verify`` / ``batch_verify`` (two pairings, O(1), host) are outside the hot path
and are not rebuilt.  Polynomials are coefficient lists of Python ints (canonical Fr values), low
degree first, as `DensePolynomial::coeffs`.

`open` runs entirely on the device (zkp_kzg_open: chunked Horner, division by X - z as a weighted suffix
scan, commitment of the quotient); no field or group arithmetic lives in this file.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import fields
from .fields import FR_MODULUS

Point = Optional[Tuple[int, int]]


def _trim(coeffs: Sequence[int]) -> List[int]:
    """`DensePolynomial::from_coefficients_vec` drops trailing zero coefficients."""
    c = [int(v) % FR_MODULUS for v in coeffs]
    while c and c[-1] == 0:
        c.pop()
    return c


@dataclass(frozen=True)
class KzgCommitment:
    """kzg/src/commitment.rs:5 -- newtype over the normalised affine G1 point (None = identity)."""

    point: Point

    def inner(self) -> Point:
        return self.point


@dataclass(frozen=True)
class KzgOpening:
    """kzg/src/opening.rs:12 -- (quotient commitment, evaluation)."""

    point: Point
    evaluation: int


class Srs:
    """kzg/src/srs.rs:14-21 (G1 part; the two G2 points only feed the pairing check)."""

    def __init__(self, g1_limbs: np.ndarray):
        self._g1 = np.ascontiguousarray(g1_limbs, dtype=np.uint64).reshape(-1, 12)

    @classmethod
    def new_from_secret(cls, engine, secret: int, circuit_size: int) -> "Srs":
        """srs.rs:48-69: circuit_size + 3 powers of the secret times the generator."""
        pts = engine.srs_generate(int(secret) % FR_MODULUS, circuit_size + 3, want_points=True)
        return cls(pts)

    @classmethod
    def from_points(cls, points: Sequence[Point]) -> "Srs":
        return cls(fields.g1_to_array(points))

    def g1_limbs(self) -> np.ndarray:
        return self._g1

    def g1_points(self) -> List[Point]:
        """srs.rs:78-80 (returns a copy, as the reference does)."""
        return fields.g1_from_array(self._g1)

    def __len__(self) -> int:
        return self._g1.shape[0]


class KzgScheme:
    """kzg/src/scheme.rs:22-36.  Construction uploads the SRS to HBM once; the reference instead
    clones the whole `Vec<G1Affine>` on every commit (srs.rs:78-80 at scheme.rs:85)."""

    def __init__(self, engine, srs: Srs, precompute: bool = True):
        self.engine = engine
        self.srs = srs
        engine.srs_upload(srs.g1_limbs())
        if precompute and len(srs):
            try:
                engine.srs_precompute()  # fixed-base table: one bucket set for all windows of every commit
            except Exception as ex:  # the table is an accelerator (12 x the SRS in HBM): without room for it
                if getattr(ex, "status", None) != 3:  # ZKP_B200_ERR_OOM -> the windowed GPU path serves the commits
                    raise

    # scheme.rs:84-96
    def _evaluate_in_s(self, coeffs: Sequence[int]) -> Point:
        degree = max(len(coeffs) - 1, 0)  # ark-poly: degree of the zero polynomial is 0
        if not len(self.srs) > degree:
            raise AssertionError("assertion failed: g1_points.len() > polynomial.degree()")  # scheme.rs:86
        scalars = fields.fr_to_mont_array(coeffs) if len(coeffs) else np.zeros((0, 4), dtype=np.uint64)
        out, inf = self.engine.msm(scalars)
        return None if inf else fields.g1_from_array(out)[0]

    def commit(self, polynomial: Sequence[int]) -> KzgCommitment:
        """scheme.rs:49-52 (`polynomial` is already a trimmed DensePolynomial in the reference)."""
        return KzgCommitment(self._evaluate_in_s(_trim(polynomial)))

    def commit_vector(self, coeffs: Sequence[int]) -> KzgCommitment:
        """scheme.rs:63-67."""
        return KzgCommitment(self._evaluate_in_s(_trim(coeffs)))

    def commit_para(self, para: int) -> KzgCommitment:
        """scheme.rs:78-82: para * g1_points[0]."""
        if len(self.srs) == 0:
            raise ValueError("called `Option::unwrap()` on a `None` value")
        scalars = fields.fr_to_mont_array([para])
        out, inf = self.engine.msm(scalars)
        return KzgCommitment(None if inf else fields.g1_from_array(out)[0])

    def open(self, polynomial: Sequence[int], z: int) -> KzgOpening:
        """scheme.rs:108-120: evaluation, division by (X - z) and the commitment of the quotient all on the device
        (zkp_kzg_open)."""
        coeffs = _trim(polynomial)
        if not coeffs:
            raise ValueError("at least 1")  # scheme.rs:112 expect("at least 1")
        out, inf, y = self.engine.kzg_open(fields.fr_to_mont_array(coeffs), int(z) % FR_MODULUS)
        return KzgOpening(None if inf else fields.g1_from_array(out)[0], y)

    def open_vector(self, coeffs: Sequence[int], z: int) -> KzgOpening:
        """scheme.rs:132-142."""
        return self.open(coeffs, z)
