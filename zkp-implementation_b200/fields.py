"""Host-side marshalling between Python integers and the arkworks in-memory layout.

arkworks stores ``Fr`` as ``BigInt<4>([u64; 4])`` and ``Fq`` as ``BigInt<6>`` -- little-endian limbs of
the MONTGOMERY representation (a * 2^256 mod r, a * 2^384 mod p).  The C ABI takes exactly those
limbs (kzg/src/types.rs:6-10 types), so these helpers are what a test or benchmark uses in place of
the Rust shim's zero-copy ``fr.0.0``.  Pure marshalling: no group or NTT arithmetic lives here.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

FQ_MODULUS = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_R = (1 << 256) % FR_MODULUS
FQ_R = (1 << 384) % FQ_MODULUS
FR_RINV = pow(FR_R, -1, FR_MODULUS)
FQ_RINV = pow(FQ_R, -1, FQ_MODULUS)


def _ints_to_limbs(vals: Iterable[int], nbytes: int) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, nbytes // 8).copy()


def _limbs_to_ints(arr: np.ndarray, nbytes: int) -> List[int]:
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


def fr_to_mont_array(vals: Iterable[int]) -> np.ndarray:
    return _ints_to_limbs(((int(v) % FR_MODULUS) * FR_R % FR_MODULUS for v in vals), 32)


def fr_from_mont_array(arr: np.ndarray) -> List[int]:
    return [v * FR_RINV % FR_MODULUS for v in _limbs_to_ints(arr, 32)]


def fq_to_mont_array(vals: Iterable[int]) -> np.ndarray:
    return _ints_to_limbs(((int(v) % FQ_MODULUS) * FQ_R % FQ_MODULUS for v in vals), 48)


def fq_from_mont_array(arr: np.ndarray) -> List[int]:
    return [v * FQ_RINV % FQ_MODULUS for v in _limbs_to_ints(arr, 48)]


Point = Optional[Tuple[int, int]]  # None = point at infinity


def g1_to_array(points: Sequence[Point]) -> np.ndarray:
    """Affine points -> n x 12 u64 (x || y Montgomery); infinity -> the (0, 0) sentinel."""
    flat: List[int] = []
    for pt in points:
        if pt is None:
            flat += [0, 0]
        else:
            flat += [pt[0], pt[1]]
    return fq_to_mont_array(flat).reshape(-1, 12)


def g1_from_array(arr: np.ndarray) -> List[Point]:
    vals = fq_from_mont_array(np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 6))
    out: List[Point] = []
    for i in range(0, len(vals), 2):
        x, y = vals[i], vals[i + 1]
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """`count` outputs of splitmix64 started at `seed` (vectorised; the same generator the test oracles use)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & (2**64 - 1)) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_fr_mont(seed: int, n: int) -> np.ndarray:
    """n pseudo-random Fr elements as Montgomery limbs (n x 4 u64): splitmix64 limbs with the top two
    bits cleared (< 2^254 < r, so every draw is a valid representation -- the same "limbs are the
    Montgomery form" convention as ark-ff's `Fr::rand`)."""
    a = splitmix64_stream(seed, 4 * n).reshape(n, 4).copy()
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a
