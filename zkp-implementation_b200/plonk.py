"""Host-side mirror of the reference's `plonk` crate surface for the hot path, over the C ABI in
``include/zkp_plonk.h`` (implemented by ``host/plonk.cpp`` inside ``libzkp_b200.so``).

=====================================================  =======================================
reference (plonk/src)                                  here
=====================================================  =======================================
``Circuit::default`` / ``add_*_gate``  circuit.rs:85   :class:`Circuit`
``Circuit::compile``               circuit.rs:166-197  :meth:`Circuit.compile`
``CompiledCircuit``             compiled_circuit.rs:5  :class:`CompiledCircuit`
``prover::generate_proof``          prover.rs:61-293   :func:`generate_proof`
``prover::Proof``                   prover.rs:24-58    :class:`Proof`
=====================================================  =======================================

Wires are ``(column, row, value)`` triples exactly as in the reference.  Field values are Python ints
(canonical).  The nine blinding scalars the reference draws from ``StdRng::from_entropy()``
(prover.rs:68) are an explicit argument, which is what makes proofs reproducible byte for byte.
Every G1 sum and every transform of `generate_proof` runs on the GPU engine the scheme was built on;
this file only marshals arguments.  No arithmetic fallback lives here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import fields
from .fields import FR_MODULUS

Point = Optional[Tuple[int, int]]
Wire = Tuple[int, int, int]

ERR_REMAINDER, ERR_INVALID_POSITION, ERR_TOO_FEW_GATES, ERR_TRANSCRIPT = 20, 21, 22, 23

_vp = ctypes.c_void_p
PLONK_ABI = {
    "zkp_plonk_circuit_new": (_vp, []),
    "zkp_plonk_circuit_free": (None, [_vp]),
    "zkp_plonk_circuit_add_gates": (ctypes.c_int, [_vp, ctypes.c_size_t, _vp, _vp, _vp, _vp]),
    "zkp_plonk_circuit_len": (ctypes.c_size_t, [_vp]),
    "zkp_plonk_compile": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "zkp_plonk_compiled_free": (None, [_vp]),
    "zkp_plonk_compiled_size": (ctypes.c_size_t, [_vp]),
    "zkp_plonk_compiled_poly": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "zkp_plonk_preprocess": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]),
    "zkp_plonk_prove": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "zkp_plonk_prove_products": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "zkp_plonk_prove_reference": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "zkp_plonk_prove_sharded": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t,
                                               ctypes.c_size_t, _vp, _vp]),
    "zkp_plonk_numden_dev": (ctypes.c_int, [_vp, _vp]),
    "zkp_plonk_quotient_dev": (ctypes.c_int, [_vp, _vp]),
    "zkp_plonk_gate_check_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]),
    # Fiat-Shamir transcript pieces (known-answer tests)
    "zkp_transcript_sha256": (None, [_vp, ctypes.c_size_t, _vp]),
    "zkp_transcript_pcg32_output": (ctypes.c_uint32, [ctypes.c_uint64]),
    "zkp_transcript_seed_from_u64": (None, [ctypes.c_uint64, _vp]),
    "zkp_transcript_chacha_words": (None, [_vp, ctypes.c_int, ctypes.c_size_t, _vp]),
    "zkp_transcript_g1_serialize": (None, [_vp, _vp]),
    "zkp_transcript_challenges": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_size_t, _vp]),
}

_PANICS = {
    ERR_REMAINDER: "No remainder here",                                  # prover.rs:404/431/441
    ERR_INVALID_POSITION: "Invalid position",                             # circuit.rs:221
    ERR_TOO_FEW_GATES: "argument of integer logarithm must be positive",  # circuit.rs:151
    ERR_TRANSCRIPT: "I'm hungry! Feed me something first",                # challenge.rs:61-63
}


class PlonkPanic(RuntimeError):
    """A status where the reference panics; ``status`` is the ZKP_PLONK_ERR_* / ZKP_B200_ERR_* code."""

    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status


def _bind(lib) -> None:
    if getattr(lib, "_plonk_bound", False):
        return
    for name, (res, args) in PLONK_ABI.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._plonk_bound = True


def _raise(engine, st: int) -> None:
    if st == 0:
        return
    msg = _PANICS.get(st) or engine.lib.zkp_strerror(st).decode()
    raise PlonkPanic(st, msg)


POLY_NAMES = ("f_a", "f_b", "f_c", "q_l", "q_r", "q_o", "q_m", "q_c", "pi", "s_sigma_1", "s_sigma_2", "s_sigma_3")


class Circuit:
    """plonk/src/circuit.rs:17-115.  Gates are buffered here and handed to the native builder in bulk."""

    ADD, MUL, CONST = 0, 1, 2

    def __init__(self):
        self._kinds: List[int] = []
        self._pos: List[int] = []
        self._vals: List[int] = []
        self._pis: List[int] = []

    def _add(self, a: Wire, b: Wire, c: Wire, kind: int, pi: int) -> None:
        for w in (a, b, c):
            if w[0] < 0 or w[1] < 0:
                raise OverflowError("usize position")
            self._pos += [int(w[0]), int(w[1])]
            self._vals.append(int(w[2]) % FR_MODULUS)
        self._kinds.append(kind)
        self._pis.append(int(pi) % FR_MODULUS)

    def add_addition_gate(self, a: Wire, b: Wire, c: Wire, pi: int = 0) -> None:
        self._add(a, b, c, self.ADD, pi)

    def add_multiplication_gate(self, a: Wire, b: Wire, c: Wire, pi: int = 0) -> None:
        self._add(a, b, c, self.MUL, pi)

    def add_constant_gate(self, a: Wire, b: Wire, c: Wire, pi: int = 0) -> None:
        self._add(a, b, c, self.CONST, pi)

    def add_gates_bulk(self, kinds: np.ndarray, positions: np.ndarray, values_mont: np.ndarray, pis_mont: np.ndarray):
        """Large synthetic circuits: arrays already in ABI layout (count; count x 6; count x 3 x 4; count x 4)."""
        self._bulk = (np.ascontiguousarray(kinds, dtype=np.uint8), np.ascontiguousarray(positions, dtype=np.uint64),
                      np.ascontiguousarray(values_mont, dtype=np.uint64), np.ascontiguousarray(pis_mont, dtype=np.uint64))

    def __len__(self) -> int:
        return len(self._kinds) + (len(self._bulk[0]) if getattr(self, "_bulk", None) else 0)

    def compile(self, engine) -> "CompiledCircuit":
        """circuit.rs:166-197: pad to a power of two, interpolate the nine gate columns and the three
        sigma columns (one batched iNTT on the GPU)."""
        lib = engine.lib
        _bind(lib)
        h = lib.zkp_plonk_circuit_new()
        if not h:
            raise MemoryError
        try:
            if self._kinds:
                kinds = np.array(self._kinds, dtype=np.uint8)
                pos = np.array(self._pos, dtype=np.uint64)
                vals = fields.fr_to_mont_array(self._vals)
                pis = fields.fr_to_mont_array(self._pis)
                _raise(engine, lib.zkp_plonk_circuit_add_gates(h, len(kinds), kinds.ctypes.data, pos.ctypes.data,
                                                               vals.ctypes.data, pis.ctypes.data))
            if getattr(self, "_bulk", None):
                k, p, v, q = self._bulk
                _raise(engine, lib.zkp_plonk_circuit_add_gates(h, len(k), k.ctypes.data, p.ctypes.data, v.ctypes.data,
                                                               q.ctypes.data))
            out = _vp()
            _raise(engine, lib.zkp_plonk_compile(engine._h, h, ctypes.byref(out)))
            return CompiledCircuit(engine, out)
        finally:
            lib.zkp_plonk_circuit_free(h)


class CompiledCircuit:
    """plonk/src/compiled_circuit.rs:5-42 (gate + copy constraints as coefficient vectors)."""

    def __init__(self, engine, handle):
        self.engine = engine
        self._h = handle
        self.size = int(engine.lib.zkp_plonk_compiled_size(handle))

    def poly(self, name: str) -> List[int]:
        """Trimmed coefficients (canonical ints) of one compiled polynomial; names as in constraint.rs."""
        buf = np.zeros((self.size, 4), dtype=np.uint64)
        _raise(self.engine, self.engine.lib.zkp_plonk_compiled_poly(self._h, POLY_NAMES.index(name), buf.ctypes.data))
        c = fields.fr_from_mont_array(buf)
        while c and c[-1] == 0:
            c.pop()
        return c

    PREPROCESSED = ("q_m", "q_l", "q_r", "q_o", "q_c", "s_sigma_1", "s_sigma_2", "s_sigma_3")

    def preprocess(self, refresh: bool = False) -> dict:
        """`get_circuit_commitment` (verifier.rs:160-185) / `CommonPreprocessedInput::new` (cpi_parser.rs:76-106): the eight
        selector / permutation commitments as one batched MSM, cached on the compiled circuit."""
        out = np.zeros((8, 12), dtype=np.uint64)
        _raise(self.engine, self.engine.lib.zkp_plonk_preprocess(self.engine._h, self._h, out.ctypes.data, 1 if refresh else 0))
        return dict(zip(self.PREPROCESSED, fields.g1_from_array(out)))

    def close(self) -> None:
        if self._h:
            # the native object frees its device buffers through the context: only while the engine is still open
            if getattr(self.engine, "_h", None):
                self.engine.lib.zkp_plonk_compiled_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ProofStruct(ctypes.Structure):
    _fields_ = [("commitments", ctypes.c_uint64 * 12 * 9), ("evaluations", ctypes.c_uint64 * 4 * 6),
                ("u", ctypes.c_uint64 * 4), ("degree", ctypes.c_uint64)]


@dataclass
class Proof:
    """plonk/src/prover.rs:24-58."""

    a_commit: Point
    b_commit: Point
    c_commit: Point
    z_commit: Point
    t_lo_commit: Point
    t_mid_commit: Point
    t_hi_commit: Point
    w_ev_x_commit: Point
    w_ev_wx_commit: Point
    bar_a: int
    bar_b: int
    bar_c: int
    bar_s_sigma_1: int
    bar_s_sigma_2: int
    bar_z_w: int
    u: int
    degree: int
    timings_ms: Optional[dict] = None

    def commitments(self) -> List[Point]:
        return [self.a_commit, self.b_commit, self.c_commit, self.z_commit, self.t_lo_commit, self.t_mid_commit,
                self.t_hi_commit, self.w_ev_x_commit, self.w_ev_wx_commit]

    def scalars(self) -> List[int]:
        return [self.bar_a, self.bar_b, self.bar_c, self.bar_s_sigma_1, self.bar_s_sigma_2, self.bar_z_w]

    def to_bytes(self) -> bytes:
        """Canonical byte string of the proof for byte-for-byte comparison: the nine commitments in
        ark-serialize's uncompressed G1 layout (what the transcript hashes, challenge.rs:52-55), the six
        evaluations and u as 32-byte little-endian canonical Fr, the degree as u64 LE."""
        out = b""
        for p in self.commitments():
            out += (bytes([0x40]) + bytes(95)) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
        for s in self.scalars() + [self.u]:
            out += int(s).to_bytes(32, "little")
        return out + int(self.degree).to_bytes(8, "little")


def generate_proof(compiled_circuit: CompiledCircuit, blinding: Sequence[int], products: bool = False,
                   reference_acc: bool = False, timings: bool = True) -> Proof:
    """prover.rs:61-293 against the SRS resident on the compiled circuit's engine (`KzgScheme(engine, srs)`
    uploads it).  ``blinding`` = b1..b9.  ``products=True`` runs the cross-check prover that performs every
    polynomial product of prover.rs separately (zkp_plonk_prove_products); ``reference_acc=True`` additionally keeps
    the reference's O(n^2) `compute_acc` (zkp_plonk_prove_reference); all return the same bytes.
    ``timings=False`` passes no timing buffer: the prover then never synchronises the stream at phase boundaries."""
    eng = compiled_circuit.engine
    _bind(eng.lib)
    if len(blinding) != 9:
        raise ValueError("blinding must hold b1..b9")
    b = fields.fr_to_mont_array(blinding)
    ps = _ProofStruct()
    tm = (ctypes.c_double * 4)()
    fn = eng.lib.zkp_plonk_prove_reference if reference_acc else (
        eng.lib.zkp_plonk_prove_products if products else eng.lib.zkp_plonk_prove)
    _raise(eng, fn(eng._h, compiled_circuit._h, b.ctypes.data, ctypes.addressof(ps),
                   ctypes.addressof(tm) if timings else None))
    cm = fields.g1_from_array(np.frombuffer(bytes(ps.commitments), dtype=np.uint64).reshape(9, 12))
    ev = fields.fr_from_mont_array(np.frombuffer(bytes(ps.evaluations), dtype=np.uint64).reshape(6, 4))
    u = fields.fr_from_mont_array(np.frombuffer(bytes(ps.u), dtype=np.uint64).reshape(1, 4))[0]
    return Proof(*cm, *ev, u, int(ps.degree),
                 timings_ms={"total": tm[0], "msm": tm[1], "ntt": tm[2], "other": tm[3]})


def chain_circuit(n_gates: int, seed: int) -> Circuit:
    """Synthetic workload of SURVEY.md 8d config 4: n_gates alternating multiplication / addition gates,
    the output of gate i wired into the left input of gate i + 1 (a 2-cycle in the permutation), random
    right inputs.  Deterministic in (n_gates, seed).  Built directly in ABI layout (Montgomery limbs)."""
    import random

    rng = random.Random(seed)
    R = FR_MODULUS
    vals: List[int] = []
    a = rng.randrange(R)
    for i in range(n_gates):
        b = rng.getrandbits(254)
        out = a * b % R if i % 2 == 0 else (a + b) % R
        vals += [a, b, out]
        a = out
    idx = np.arange(n_gates, dtype=np.uint64)
    pos = np.zeros((n_gates, 6), dtype=np.uint64)
    pos[:, 0] = 2                      # a_i  -> slot of c_{i-1}
    pos[:, 1] = idx - np.uint64(1)
    pos[0, 0], pos[0, 1] = 0, 0        # first gate: fixed point
    pos[:, 2], pos[:, 3] = 1, idx      # b_i: fixed point
    pos[:, 4] = 0                      # c_i  -> slot of a_{i+1}
    pos[:, 5] = idx + np.uint64(1)
    pos[-1, 4], pos[-1, 5] = 2, n_gates - 1
    kinds = np.where(idx % 2 == 0, Circuit.MUL, Circuit.ADD).astype(np.uint8)
    c = Circuit()
    c.add_gates_bulk(kinds, pos, fields.fr_to_mont_array(vals).reshape(n_gates, 3, 4),
                     np.zeros((n_gates, 4), dtype=np.uint64))
    return c


ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


def generate_proof_sharded(compiled_circuit: CompiledCircuit, blinding: Sequence[int], rank: int, world: int,
                           srs_first: int, srs_total: int, group=None, device=None) -> Proof:
    """generate_proof with the commitments point-range-sharded over the ranks of a torch.distributed group
    (zkp_plonk_prove_sharded): this rank's engine holds SRS points [srs_first, srs_first + engine.srs_len())."""
    import torch
    import torch.distributed as dist

    eng = compiled_circuit.engine
    _bind(eng.lib)

    def _allgather(_user, send, nbytes, recv):
        try:
            buf = (ctypes.c_uint8 * nbytes).from_address(send)
            t = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
            if device is not None:
                t = t.to(device)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t, group=group)
            flat = torch.cat(outs).cpu().numpy().tobytes()
            ctypes.memmove(recv, flat, nbytes * world)
            return 0
        except Exception:  # a failed collective must not unwind through the C frames
            return 2

    cb = ALLGATHER_FN(_allgather)
    b = fields.fr_to_mont_array(blinding)
    ps = _ProofStruct()
    tm = (ctypes.c_double * 4)()
    _raise(eng, eng.lib.zkp_plonk_prove_sharded(eng._h, compiled_circuit._h, b.ctypes.data, ctypes.addressof(ps),
                                                ctypes.addressof(tm), rank, world, srs_first, srs_total,
                                                ctypes.cast(cb, ctypes.c_void_p), None))
    cm = fields.g1_from_array(np.frombuffer(bytes(ps.commitments), dtype=np.uint64).reshape(9, 12))
    ev = fields.fr_from_mont_array(np.frombuffer(bytes(ps.evaluations), dtype=np.uint64).reshape(6, 4))
    u = fields.fr_from_mont_array(np.frombuffer(bytes(ps.u), dtype=np.uint64).reshape(1, 4))[0]
    return Proof(*cm, *ev, u, int(ps.degree), timings_ms={"total": tm[0], "msm": tm[1], "ntt": tm[2], "other": tm[3]})
