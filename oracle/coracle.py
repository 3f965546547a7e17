"""ctypes wrapper over oracle/zkp_oracle.c (CPU restatement of the reference path; test infrastructure).

All arrays are numpy uint64 in the arkworks layout (Montgomery limbs): Fr = 4 limbs, G1 affine = 12
limbs x || y with (0, 0) as the point at infinity -- the same layout the engine's C ABI takes, so a
parity test feeds identical buffers to both sides.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libzkp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "zkp_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_g1_on_curve.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return ctypes.c_void_p(0 if a is None else a.ctypes.data)


def _c(a, cols: int) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, cols)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def msm_naive(scalars: np.ndarray, bases: np.ndarray) -> np.ndarray:
    """kzg/src/scheme.rs:84-96 verbatim (zip-truncating, per-term into_affine)."""
    s, b = _c(scalars, 4), _c(bases, 12)
    n = min(s.shape[0], b.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_msm_naive(_p(s), _p(b), ctypes.c_size_t(n), _p(out))
    return out


def msm_pippenger(scalars: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    s, b = _c(scalars, 4), _c(bases, 12)
    n = min(s.shape[0], b.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_msm_pippenger(_p(s), _p(b), ctypes.c_size_t(n), _p(out), ctypes.c_int(threads))
    return out


def srs(secret_mont: np.ndarray, count: int) -> np.ndarray:
    out = np.zeros((count, 12), dtype=np.uint64)
    sec = _c(secret_mont, 4)
    lib().orc_srs(_p(sec), ctypes.c_size_t(count), _p(out))
    return out


def on_curve(xy: np.ndarray) -> bool:
    pts = _c(xy, 12)
    return all(lib().orc_g1_on_curve(ctypes.c_void_p(pts[i].ctypes.data)) for i in range(pts.shape[0]))


def open_quotient(coeffs: np.ndarray, z_mont: np.ndarray):
    c = _c(coeffs, 4)
    n = c.shape[0]
    q = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
    y = np.zeros(4, dtype=np.uint64)
    z = _c(z_mont, 4)
    lib().orc_open_quotient(_p(c), ctypes.c_size_t(n), _p(z), _p(q), _p(y))
    return q, y


def ntt(data: np.ndarray, log_n: int, inverse: bool = False, coset_mont: Optional[np.ndarray] = None,
        threads: int = 0) -> np.ndarray:
    """Returns a transformed COPY (one polynomial of 2^log_n elements)."""
    d = _c(data, 4).copy()
    assert d.shape[0] == 1 << log_n
    cs = None if coset_mont is None else _c(coset_mont, 4)
    lib().orc_ntt(_p(d), ctypes.c_int(log_n), ctypes.c_int(1 if inverse else 0), _p(cs), ctypes.c_int(threads))
    return d


def poly_mul(a: np.ndarray, b: np.ndarray, threads: int = 0) -> np.ndarray:
    a, b = _c(a, 4), _c(b, 4)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((0, 4), dtype=np.uint64)
    out = np.zeros((a.shape[0] + b.shape[0] - 1, 4), dtype=np.uint64)
    lib().orc_poly_mul(_p(a), ctypes.c_size_t(a.shape[0]), _p(b), ctypes.c_size_t(b.shape[0]), _p(out),
                       ctypes.c_int(threads))
    return out
