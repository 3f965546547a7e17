"""ctypes handle over oracle/_build/libzkp_cpu_ref.so: the product's host orchestration of the PLONK prover
(zkp-implementation_b200/host/plonk.cpp, unchanged) linked against a CPU backend of the C ABI (oracle/cpu_backend.cpp,
kernels from oracle/zkp_oracle.c).  TEST / BASELINE INFRASTRUCTURE: the object quacks like `Engine` just enough for
`zkp_implementation_b200.plonk.Circuit.compile` / `generate_proof(..., products=True | reference_acc=True)`, so the
reference's own prover algorithm can be run -- and timed -- on host cores:

    eng = CpuEngine(threads=1)                       # single-threaded, per-term MSM: the reference as written
    eng.srs_from_secret(secret, n + 3)
    proof = plonk.generate_proof(circuit.compile(eng), blinding, reference_acc=True)   # O(n^2) compute_acc kept
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import coracle

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libzkp_cpu_ref.so")


def build() -> str:
    coracle.build()
    deps = [os.path.join(HERE, "cpu_backend.cpp"), os.path.join(HERE, "zkp_oracle.c"),
            os.path.join(HERE, "..", "zkp-implementation_b200", "host", "plonk.cpp"),
            os.path.join(HERE, "..", "zkp-implementation_b200", "host", "transcript.hpp")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        import subprocess

        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


class CpuEngine:
    def __init__(self, threads: int = 1, pippenger: bool = False):
        self.lib = ctypes.CDLL(build())
        self.lib.zkp_strerror.restype = ctypes.c_char_p
        self.lib.zkp_srs_len.restype = ctypes.c_size_t
        h = ctypes.c_void_p()
        assert self.lib.zkp_cpu_ctx_create(ctypes.byref(h), int(threads), 1 if pippenger else 0) == 0
        self._h = h
        self.threads, self.pippenger = threads, pippenger

    def srs_upload(self, xy: np.ndarray) -> None:
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 12)
        assert self.lib.zkp_srs_upload(self._h, ctypes.c_void_p(xy.ctypes.data), None, ctypes.c_size_t(xy.shape[0])) == 0

    def srs_from_secret(self, secret_mont: np.ndarray, count: int) -> None:
        """kzg/src/srs.rs:48-69 (serial scalar multiplications: use srs_upload for large counts)."""
        sec = np.ascontiguousarray(secret_mont, dtype=np.uint64).reshape(4)
        assert self.lib.zkp_srs_generate(self._h, ctypes.c_void_p(sec.ctypes.data), ctypes.c_size_t(count), None) == 0

    def srs_len(self) -> int:
        return int(self.lib.zkp_srs_len(self._h))

    def close(self) -> None:
        if self._h:
            self.lib.zkp_cpu_ctx_destroy(self._h)
            self._h = None
