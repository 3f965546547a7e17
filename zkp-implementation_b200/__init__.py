"""zkp-implementation_b200 -- B200-native G1 MSM + Fr NTT engine behind the reference's kzg/plonk seam.

This package is a thin ctypes binding over the C ABI in ``include/zkp_b200.h`` (implemented by
``csrc/*.cu`` -> ``libzkp_b200.so``).  It is plumbing for tests, ``bench.py`` and
``torch.distributed`` runs; the product is the shared library.

There is NO CPU fallback: importing works anywhere (so the CPU test-suite can check symbols and
host logic), but creating an :class:`Engine` without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from . import fields
from .fields import FR_MODULUS, FQ_MODULUS

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libzkp_b200.so"

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)

# name -> (restype, argtypes); every symbol include/zkp_b200.h declares
ABI = {
    "zkp_ctx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "zkp_ctx_create_multi": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int]),
    "zkp_ctx_shards": (ctypes.c_int, [ctypes.c_void_p]),
    "zkp_ctx_destroy": (None, [ctypes.c_void_p]),
    "zkp_ctx_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "zkp_ctx_set_msm_window": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32]),
    "zkp_ctx_set_msm_affine": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "zkp_ctx_last_launches": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "zkp_ctx_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "zkp_ctx_last_phase_ms": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_int]),
    "zkp_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "zkp_srs_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_srs_upload_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_srs_len": (ctypes.c_size_t, [ctypes.c_void_p]),
    "zkp_srs_precompute": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32]),
    "zkp_srs_generate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "zkp_msm_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_msm_g1_bases": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_msm_g1_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_msm_g1_multi_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_msm_g1_multi_partial_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p]),
    "zkp_srs_generate_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                              ctypes.c_void_p]),
    "zkp_msm_g1_partial_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_void_p]),
    "zkp_g1_fold_partials": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_ntt_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_int,
                                  ctypes.c_void_p]),
    "zkp_ntt_fr_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_int,
                                      ctypes.c_void_p]),
    "zkp_poly_mul_fr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_void_p]),
    "zkp_fr_mul_pointwise_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_kzg_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_ntt_dist_rows_log": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32]),
    "zkp_ntt_dist_stage_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_ntt_dist_permute_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                                ctypes.c_uint32, ctypes.c_int]),
    "zkp_dev_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "zkp_dev_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_dev_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_ipc_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_ipc_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "zkp_ipc_close": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_dev_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_dev_download": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_dev_zero": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_fr_powers_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_fr_batch_inverse_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_fr_scan_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]),
    "zkp_fr_lincomb_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_fr_add_at_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "zkp_fr_eval_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "zkp_fr_trimmed_len_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.POINTER(ctypes.c_size_t)]),
    "zkp_g1_mul_srs0": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]),
    "zkp_sort_pairs_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                          ctypes.c_int]),
    "zkp_scan_exclusive_u32_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "zkp_g1_generate_bases_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p]),
    "zkp_g1_generate_bases_range_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_size_t,
                                                       ctypes.c_void_p]),
    "zkp_bench_imad_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                           ctypes.POINTER(ctypes.c_double)]),
}


class ZkpError(RuntimeError):
    """Non-zero status from the C ABI (the Rust shim maps these to panic!, as the reference does)."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"zkp_b200 status {status}: {msg}")
        self.status = status


def library_path() -> str:
    return os.path.join(PKG_DIR, LIB_NAME)


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load the C-ABI library and bind every declared symbol.  Fails loudly when it is missing."""
    path = path or library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python zkp-implementation_b200/build.py cuda` "
            "(nvcc, sm_100a).  This engine has no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def _ptr(a) -> ctypes.c_void_p:
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (device or host)
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(int(a))


class Engine:
    """One context = one GPU + one CUDA stream (``zkp_ctx``)."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None, stream: Optional[int] = None,
                 devices: Optional[Sequence[int]] = None):
        """``devices``: single-process multi-GPU context (zkp_ctx_create_multi): the resident SRS is sharded by point range
        over these devices and every commitment against it runs on all of them; everything else runs on devices[0]."""
        self.lib = load_library(lib_path)
        h = ctypes.c_void_p()
        if devices is not None:
            ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            st = self.lib.zkp_ctx_create_multi(ctypes.byref(h), ctypes.cast(ids, ctypes.c_void_p), len(devices))
            device = int(devices[0])
        else:
            st = self.lib.zkp_ctx_create(ctypes.byref(h), int(device))
        self._h = h if st == 0 else None
        self._check(st)
        self.device = device
        if stream is not None:
            self._check(self.lib.zkp_ctx_set_stream(self._h, ctypes.c_void_p(stream)))

    # -- plumbing ---------------------------------------------------------------------------------
    def _check(self, st: int) -> None:
        if st != 0:
            raise ZkpError(st, self.lib.zkp_strerror(st).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.zkp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self) -> int:
        return int(self.lib.zkp_ctx_shards(self._h))

    def is_cuda(self) -> bool:
        """False for the CPU kernel emulator of the test-suite (tests/emu)."""
        return os.path.basename(getattr(self.lib, "_name", "")) == LIB_NAME

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.zkp_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def set_msm_window(self, bits: int) -> None:
        self._check(self.lib.zkp_ctx_set_msm_window(self._h, bits))

    def set_msm_affine(self, rounds: int) -> None:
        """Batched-affine tree rounds before the XYZZ finish (-1 = automatic, 0 = off)."""
        self._check(self.lib.zkp_ctx_set_msm_affine(self._h, rounds))

    def last_affine_rounds(self) -> int:
        return int(self.lib.zkp_ctx_last_launches(self._h, 4))

    def last_launches(self, kind: str) -> int:
        return int(self.lib.zkp_ctx_last_launches(self._h, 0 if kind == "msm" else 1))

    def last_msm_shape(self) -> Tuple[int, int]:
        """(window bits c, number of windows W) used by the last MSM."""
        return int(self.lib.zkp_ctx_last_launches(self._h, 2)), int(self.lib.zkp_ctx_last_launches(self._h, 3))

    PHASES = ("recode", "sort", "bounds_tasks", "accumulate", "reduce")

    def set_profiling(self, on: bool) -> None:
        self._check(self.lib.zkp_ctx_set_profiling(self._h, 1 if on else 0))

    def last_phase_ms(self) -> dict:
        return {name: float(self.lib.zkp_ctx_last_phase_ms(self._h, i)) for i, name in enumerate(self.PHASES)}

    def last_affine_profile(self) -> dict:
        """First-round addition kernel time and the number of points entering / leaving each batched-affine round of
        the last MSM (profiling on)."""
        r = self.last_affine_rounds()
        pts = [int(self.lib.zkp_ctx_last_phase_ms(self._h, 200 + i)) for i in range(min(r, 7) + 1)] if r else []
        return {"rounds": r, "add_kernel_round1_ms": float(self.lib.zkp_ctx_last_phase_ms(self._h, 100)), "points": pts}

    # -- SRS --------------------------------------------------------------------------------------
    def srs_upload(self, xy: np.ndarray, infinity: Optional[np.ndarray] = None) -> None:
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 12)
        inf = None if infinity is None else np.ascontiguousarray(infinity, dtype=np.uint8)
        self._check(self.lib.zkp_srs_upload(self._h, _ptr(xy), _ptr(inf), xy.shape[0]))

    def srs_upload_dev(self, bases_dev, n: int) -> None:
        self._check(self.lib.zkp_srs_upload_dev(self._h, _ptr(bases_dev), n))

    def srs_precompute(self, window_bits: int = 0) -> None:
        """Build the fixed-base table over the resident SRS (one bucket set for all windows)."""
        self._check(self.lib.zkp_srs_precompute(self._h, window_bits))

    def srs_len(self) -> int:
        return int(self.lib.zkp_srs_len(self._h))

    def srs_generate(self, secret: int, n: int, want_points: bool = True, first: int = 0) -> Optional[np.ndarray]:
        """Resident SRS = [secret^i G] for i in [first, first + n) (first > 0: one rank's shard of a sharded SRS)."""
        sec = fields.fr_to_mont_array([secret])
        out = np.zeros((n, 12), dtype=np.uint64) if want_points else None
        self._check(self.lib.zkp_srs_generate_range(self._h, _ptr(sec), first, n, _ptr(out)))
        return out

    # -- MSM --------------------------------------------------------------------------------------
    def msm(self, scalars: np.ndarray, bases: Optional[np.ndarray] = None,
            infinity: Optional[np.ndarray] = None) -> Tuple[np.ndarray, bool]:
        """Host buffers in, normalised affine point out (x || y Montgomery limbs, infinity flag)."""
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(12, dtype=np.uint64)
        inf = ctypes.c_uint8(0)
        if bases is None:
            st = self.lib.zkp_msm_g1(self._h, _ptr(scalars), scalars.shape[0], _ptr(out), ctypes.byref(inf))
        else:
            bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 12)
            assert bases.shape[0] == scalars.shape[0]
            finf = None if infinity is None else np.ascontiguousarray(infinity, dtype=np.uint8)
            st = self.lib.zkp_msm_g1_bases(self._h, _ptr(scalars), _ptr(bases), _ptr(finf), scalars.shape[0], _ptr(out),
                                           ctypes.byref(inf))
        self._check(st)
        return out, bool(inf.value)

    def msm_dev(self, scalars_dev, bases_dev, n: int) -> Tuple[np.ndarray, bool]:
        out = np.zeros(12, dtype=np.uint64)
        inf = ctypes.c_uint8(0)
        self._check(self.lib.zkp_msm_g1_dev(self._h, _ptr(scalars_dev), _ptr(bases_dev), n, _ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_multi_dev(self, scalars_devs: Sequence, lens: Sequence[int]):
        """Several commitments against the resident SRS as one pipeline -> list of (xy limbs, infinity)."""
        k = len(lens)
        ptrs = (ctypes.c_void_p * k)(*[_ptr(s) for s in scalars_devs])
        ln = (ctypes.c_size_t * k)(*lens)
        out = np.zeros((k, 12), dtype=np.uint64)
        inf = np.zeros(k, dtype=np.uint8)
        self._check(self.lib.zkp_msm_g1_multi_dev(self._h, k, ctypes.cast(ptrs, ctypes.c_void_p),
                                                  ctypes.cast(ln, ctypes.c_void_p), _ptr(out), _ptr(inf)))
        return [(out[j], bool(inf[j])) for j in range(k)]

    def kzg_open(self, coeffs: np.ndarray, z: int):
        """`KzgScheme::open` on the device: (witness xy limbs, infinity flag, evaluation as a canonical int)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        za = fields.fr_to_mont_array([z])
        out = np.zeros(12, dtype=np.uint64)
        y = np.zeros((1, 4), dtype=np.uint64)
        inf = ctypes.c_uint8(0)
        self._check(self.lib.zkp_kzg_open(self._h, _ptr(coeffs), coeffs.shape[0], _ptr(za), _ptr(out), ctypes.byref(inf),
                                          _ptr(y)))
        return out, bool(inf.value), fields.fr_from_mont_array(y)[0]

    def msm_partial_dev(self, scalars_dev, bases_dev, n: int) -> np.ndarray:
        out = np.zeros(24, dtype=np.uint64)
        self._check(self.lib.zkp_msm_g1_partial_dev(self._h, _ptr(scalars_dev), _ptr(bases_dev), n, _ptr(out)))
        return out

    def fold_partials(self, partials: np.ndarray) -> Tuple[np.ndarray, bool]:
        partials = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, 24)
        out = np.zeros(12, dtype=np.uint64)
        inf = ctypes.c_uint8(0)
        self._check(self.lib.zkp_g1_fold_partials(_ptr(partials), partials.shape[0], _ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    # -- NTT --------------------------------------------------------------------------------------
    def ntt(self, data: np.ndarray, log_n: int, batch: int = 1, inverse: bool = False,
            coset: Optional[int] = None) -> np.ndarray:
        """In-place on a host array of batch * 2^log_n Montgomery Fr elements (uint64 x 4 each)."""
        assert data.dtype == np.uint64 and data.flags.c_contiguous and data.size == (batch << log_n) * 4
        cs = None if coset is None else fields.fr_to_mont_array([coset])
        self._check(self.lib.zkp_ntt_fr(self._h, _ptr(data), log_n, batch, 1 if inverse else 0, _ptr(cs)))
        return data

    def ntt_dev(self, data_dev, log_n: int, batch: int = 1, inverse: bool = False, coset: Optional[int] = None) -> None:
        cs = None if coset is None else fields.fr_to_mont_array([coset])
        self._check(self.lib.zkp_ntt_fr_dev(self._h, _ptr(data_dev), log_n, batch, 1 if inverse else 0, _ptr(cs)))

    def poly_mul(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        la, lb = a.shape[0], b.shape[0]
        if la == 0 or lb == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        out = np.zeros((la + lb - 1, 4), dtype=np.uint64)
        self._check(self.lib.zkp_poly_mul_fr(self._h, _ptr(a), la, _ptr(b), lb, _ptr(out)))
        return out

    def mul_pointwise_dev(self, a_dev, b_dev, n: int) -> None:
        self._check(self.lib.zkp_fr_mul_pointwise_dev(self._h, _ptr(a_dev), _ptr(b_dev), n))

    # -- multi-GPU four-step NTT (see dist.DistNtt) ------------------------------------------------
    def ntt_dist_rows_log(self, log_n: int, world: int) -> int:
        return int(self.lib.zkp_ntt_dist_rows_log(log_n, world))

    def ntt_dist_stage_dev(self, data_dev, log_n: int, rank: int, world: int, inverse: bool = False,
                           coset: Optional[int] = None, peers: Optional[Sequence[int]] = None) -> None:
        cs = None if coset is None else fields.fr_to_mont_array([coset])
        arr = None
        if peers is not None:
            arr = (ctypes.c_void_p * len(peers))(*[ctypes.c_void_p(int(p)) for p in peers])
        self._check(self.lib.zkp_ntt_dist_stage_dev(self._h, _ptr(data_dev), log_n, rank, world, 1 if inverse else 0,
                                                    _ptr(cs), ctypes.cast(arr, ctypes.c_void_p) if arr is not None else None))

    def ntt_dist_permute_dev(self, in_dev, out_dev, log_n: int, world: int, inverse: bool = False) -> None:
        self._check(self.lib.zkp_ntt_dist_permute_dev(self._h, _ptr(in_dev), _ptr(out_dev), log_n, world,
                                                      1 if inverse else 0))

    def dev_alloc(self, nbytes: int) -> int:
        p = ctypes.c_void_p()
        self._check(self.lib.zkp_dev_alloc(self._h, nbytes, ctypes.byref(p)))
        return int(p.value or 0)

    def dev_free(self, ptr: int) -> None:
        self._check(self.lib.zkp_dev_free(self._h, ctypes.c_void_p(ptr)))

    def dev_copy(self, dst, src, nbytes: int) -> None:
        self._check(self.lib.zkp_dev_copy(self._h, _ptr(dst), _ptr(src), nbytes))

    def ipc_export(self, ptr: int) -> bytes:
        buf = (ctypes.c_uint8 * 64)()
        self._check(self.lib.zkp_ipc_export(self._h, ctypes.c_void_p(ptr), ctypes.cast(buf, ctypes.c_void_p)))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (ctypes.c_uint8 * 64)(*handle)
        p = ctypes.c_void_p()
        self._check(self.lib.zkp_ipc_open(self._h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(p)))
        return int(p.value or 0)

    def ipc_close(self, ptr: int) -> None:
        self._check(self.lib.zkp_ipc_close(self._h, ctypes.c_void_p(ptr)))

    # -- device-resident Fr vectors (csrc/poly.cu) -------------------------------------------------
    def vec(self, data: Optional[np.ndarray] = None, n: Optional[int] = None) -> "DevVec":
        """Device vector of Fr elements, from an (n, 4) uint64 Montgomery array or zero-filled."""
        return DevVec(self, data=data, n=n)

    def fr_powers(self, out: "DevVec", base: int, first: int = 1) -> None:
        b, f = fields.fr_to_mont_array([base]), fields.fr_to_mont_array([first])
        self._check(self.lib.zkp_fr_powers_dev(self._h, _ptr(out.ptr), _ptr(b), _ptr(f), out.n))

    def fr_batch_inverse(self, v: "DevVec") -> None:
        self._check(self.lib.zkp_fr_batch_inverse_dev(self._h, _ptr(v.ptr), v.n))

    def fr_scan(self, v: "DevVec", op: str = "mul", reverse: bool = False) -> None:
        self._check(self.lib.zkp_fr_scan_dev(self._h, _ptr(v.ptr), v.n, 0 if op == "mul" else 1, 1 if reverse else 0))

    def fr_lincomb(self, out: "DevVec", polys: Sequence["DevVec"], coefs: Sequence[int], c0: Optional[int] = None) -> None:
        k = len(polys)
        ptrs = (ctypes.c_void_p * k)(*[ctypes.c_void_p(p.ptr) for p in polys])
        lens = (ctypes.c_size_t * k)(*[p.n for p in polys])
        cf = fields.fr_to_mont_array(coefs) if k else np.zeros((0, 4), dtype=np.uint64)
        c0a = None if c0 is None else fields.fr_to_mont_array([c0])
        self._check(self.lib.zkp_fr_lincomb_dev(self._h, _ptr(out.ptr), out.n, k, ctypes.cast(ptrs, ctypes.c_void_p),
                                                ctypes.cast(lens, ctypes.c_void_p), _ptr(cf), _ptr(c0a)))

    def fr_add_at(self, v: "DevVec", idx: Sequence[int], vals: Sequence[int]) -> None:
        k = len(idx)
        ia = (ctypes.c_size_t * k)(*idx)
        va = fields.fr_to_mont_array(vals)
        self._check(self.lib.zkp_fr_add_at_dev(self._h, _ptr(v.ptr), k, ctypes.cast(ia, ctypes.c_void_p), _ptr(va)))

    def fr_eval(self, polys: Sequence["DevVec"], xs: Sequence[int]) -> list:
        k = len(polys)
        ptrs = (ctypes.c_void_p * k)(*[ctypes.c_void_p(p.ptr) for p in polys])
        lens = (ctypes.c_size_t * k)(*[p.n for p in polys])
        xa = fields.fr_to_mont_array(xs)
        out = np.zeros((k, 4), dtype=np.uint64)
        self._check(self.lib.zkp_fr_eval_dev(self._h, k, ctypes.cast(ptrs, ctypes.c_void_p),
                                             ctypes.cast(lens, ctypes.c_void_p), _ptr(xa), _ptr(out)))
        return fields.fr_from_mont_array(out)

    def fr_trimmed_len(self, v: "DevVec") -> int:
        out = ctypes.c_size_t(0)
        self._check(self.lib.zkp_fr_trimmed_len_dev(self._h, _ptr(v.ptr), v.n, ctypes.byref(out)))
        return int(out.value)

    def g1_mul_srs0(self, scalars: Sequence[int]) -> list:
        """`KzgScheme::commit_para` for several scalars at once (kzg/src/scheme.rs:78-82)."""
        sa = fields.fr_to_mont_array(scalars)
        out = np.zeros((len(scalars), 12), dtype=np.uint64)
        self._check(self.lib.zkp_g1_mul_srs0(self._h, _ptr(sa), len(scalars), _ptr(out)))
        return fields.g1_from_array(out)

    # -- radix sort / scan (csrc/sort.cu) -----------------------------------------------------------
    def sort_pairs_dev(self, keys_dev, vals_dev, n: int, key_bits: int, descending: bool = False) -> None:
        self._check(self.lib.zkp_sort_pairs_dev(self._h, _ptr(keys_dev), _ptr(vals_dev), n, key_bits, 1 if descending else 0))

    def scan_exclusive_u32_dev(self, in_dev, out_dev, n: int) -> None:
        self._check(self.lib.zkp_scan_exclusive_u32_dev(self._h, _ptr(in_dev), _ptr(out_dev), n))

    # -- synthetic workloads / microbenchmarks ----------------------------------------------------
    def generate_bases_dev(self, seed: int, n: int, bases_dev, first: int = 0) -> None:
        self._check(self.lib.zkp_g1_generate_bases_range_dev(self._h, seed & (2**64 - 1), first, n, _ptr(bases_dev)))

    def imad_peak(self) -> Tuple[float, float]:
        w, l = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self.lib.zkp_bench_imad_peak(self._h, ctypes.byref(w), ctypes.byref(l)))
        return w.value, l.value


class DevVec:
    """n Fr elements (32-byte Montgomery) in device memory owned through zkp_dev_alloc / zkp_dev_free."""

    def __init__(self, engine: Engine, data: Optional[np.ndarray] = None, n: Optional[int] = None):
        self.eng = engine
        if data is not None:
            data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
            n = data.shape[0]
        self.n = int(n or 0)
        self.ptr = engine.dev_alloc(max(self.n, 1) * 32)
        if data is not None and self.n:
            engine._check(engine.lib.zkp_dev_upload(engine._h, _ptr(self.ptr), _ptr(data), self.n * 32))
        elif self.n:
            engine._check(engine.lib.zkp_dev_zero(engine._h, _ptr(self.ptr), self.n * 32))

    def get(self) -> np.ndarray:
        out = np.zeros((self.n, 4), dtype=np.uint64)
        if self.n:
            self.eng._check(self.eng.lib.zkp_dev_download(self.eng._h, _ptr(out), _ptr(self.ptr), self.n * 32))
        return out

    def ints(self) -> list:
        return fields.fr_from_mont_array(self.get())

    def free(self) -> None:
        if self.ptr and getattr(self.eng, "_h", None):
            self.eng.dev_free(self.ptr)
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


from . import dist, plonk  # noqa: E402
from .kzg import KzgCommitment, KzgOpening, KzgScheme, Srs  # noqa: E402

__all__ = ["Engine", "ZkpError", "load_library", "library_path", "ABI", "fields", "plonk", "Srs", "KzgScheme",
           "KzgCommitment", "KzgOpening", "FR_MODULUS", "FQ_MODULUS"]
