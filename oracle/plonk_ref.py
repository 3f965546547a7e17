"""Big-integer restatement of the reference's PLONK prover (TEST INFRASTRUCTURE ONLY).

Follows, function by function (paths relative to sota-zk-labs/zkp-implementation):
  * ``Circuit`` / ``compile`` / ``cal_permutation``   plonk/src/circuit.rs:85-245, gate.rs:38-111
  * ``generate_proof`` and helpers                      plonk/src/prover.rs:61-581  (the O(n^2)
    ``compute_acc`` Horner loop at :314-369 is kept as written -- small n only)
  * ``SlicePoly``                                       plonk/src/slice_polynomial.rs:22-69
  * ``ChallengeGenerator``                              plonk/src/challenge.rs:49-89
  * ``verify``                                          plonk/src/verifier.rs:19-157, with the two
    pairings replaced by the equivalent G1 check under the known SRS secret
    (e(A, sG2) == e(B, G2)  <=>  s*A == B), which is available because tests use ``new_from_secret``.

Two seams the reference lacks are made explicit (SURVEY.md finding 5): the nine blinding scalars
b1..b9 (``StdRng::from_entropy()`` at prover.rs:68) are inputs.

NOT verifiable in this container (no Rust toolchain; SURVEY.md App. A.5/A.6, recalled semantics):
the byte layout of ``serialize_uncompressed`` for G1 (zcash format), ``StdRng::seed_from_u64``
(PCG32 expansion), ChaCha12 block/word order and ``Fr::rand`` (limbs taken as the Montgomery
representation, top bit cleared, rejection).  They are restated here and in the C++ prover
independently of each other; agreement between the two is what the tests pin.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from . import pyref as o

R = o.R
Point = Optional[Tuple[int, int]]


# ----------------------------------------------------------------------------- polynomials (ark-poly DensePolynomial)
def trim(c: Sequence[int]) -> List[int]:
    c = [x % R for x in c]
    while c and c[-1] == 0:
        c.pop()
    return c


def p_add(a, b):
    n = max(len(a), len(b))
    return trim([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def p_neg(a):
    return [(-x) % R for x in a]


def p_sub(a, b):
    return p_add(a, p_neg(b))


def p_scale(a, k):
    k %= R
    if not a or k == 0:
        return []
    return [x * k % R for x in a]


def p_mul(a, b):
    return o.poly_mul(a, b)


def p_eval(a, x):
    return o.poly_eval(a, x)


def mul_by_vanishing(a, n):
    """p * (X^n - 1)"""
    if not a:
        return []
    out = [0] * n + list(a)
    for i, c in enumerate(a):
        out[i] = (out[i] - c) % R
    return trim(out)


def divide_by_vanishing(a, n):
    """ark-poly DensePolynomial::divide_by_vanishing_poly -> (quotient, remainder)."""
    if len(a) < n:
        return [], list(a)
    q = list(a[n:])
    for i in range(1, len(a) // n):
        for j, c in enumerate(a[n * (i + 1):]):
            q[j] = (q[j] + c) % R
    rem = list(a[:n])
    for j, c in enumerate(q[:n]):
        rem[j] = (rem[j] + c) % R
    return trim(q), trim(rem)


def divide_linear(a, root):
    """a / (X - root): (quotient, remainder) by synthetic division."""
    if not a:
        return [], 0
    q = [0] * (len(a) - 1)
    carry = 0
    for i in range(len(a) - 1, 0, -1):
        carry = (a[i] + carry * root) % R
        q[i - 1] = carry
    return trim(q), (a[0] + carry * root) % R


def interpolate(evals, n):
    """Evaluations::from_vec_and_domain(v, domain).interpolate(): zero-pad, iFFT, trim."""
    log_n = n.bit_length() - 1
    return trim(o.intt(list(evals) + [0] * (n - len(evals)), log_n))


# ----------------------------------------------------------------------------- circuit (circuit.rs, gate.rs)
@dataclass
class Gate:
    a: Optional[Tuple[int, int]]
    b: Optional[Tuple[int, int]]
    c: Optional[Tuple[int, int]]
    q_l: int
    q_r: int
    q_o: int
    q_m: int
    q_c: int
    pi: int


class Circuit:
    ADD, MUL, CONST = 0, 1, 2

    def __init__(self):
        self.gates: List[Gate] = []
        self.vals: List[List[int]] = [[], [], []]

    def _add(self, a, b, c, kind, pi):
        self.vals[0].append(a[2] % R)
        self.vals[1].append(b[2] % R)
        self.vals[2].append(c[2] % R)
        pos = ((a[0], a[1]), (b[0], b[1]), (c[0], c[1]))
        npi = (-pi) % R
        if kind == self.ADD:      # gate.rs:38-56
            g = Gate(*pos, q_l=1, q_r=1, q_o=R - 1, q_m=0, q_c=0, pi=npi)
        elif kind == self.MUL:    # gate.rs:58-76
            g = Gate(*pos, q_l=0, q_r=0, q_o=R - 1, q_m=1, q_c=0, pi=npi)
        else:                     # gate.rs:78-97
            g = Gate(*pos, q_l=1, q_r=0, q_o=0, q_m=0, q_c=(-a[2]) % R, pi=npi)
        self.gates.append(g)

    def add_addition_gate(self, a, b, c, pi=0):
        self._add(a, b, c, self.ADD, pi)

    def add_multiplication_gate(self, a, b, c, pi=0):
        self._add(a, b, c, self.MUL, pi)

    def add_constant_gate(self, a, b, c, pi=0):
        self._add(a, b, c, self.CONST, pi)

    def compile(self) -> "CompiledCircuit":
        ln = len(self.gates)
        if ln == 0:
            raise ValueError("attempt to subtract with overflow")          # (len - 1) on usize
        if ln == 1:
            raise ValueError("argument of integer logarithm must be positive")  # ilog2(0), circuit.rs:151
        n = 1 << ((ln - 1).bit_length())                                    # circuit.rs:148-157
        real = range(ln)                                                    # dummy gates are appended, then skipped
        cols = {
            "a": [self.vals[0][i] for i in real], "b": [self.vals[1][i] for i in real],
            "c": [self.vals[2][i] for i in real],
            "ql": [self.gates[i].q_l for i in real], "qr": [self.gates[i].q_r for i in real],
            "qo": [self.gates[i].q_o for i in real], "qm": [self.gates[i].q_m for i in real],
            "qc": [self.gates[i].q_c for i in real], "pi": [self.gates[i].pi for i in real],
        }
        polys = {k: interpolate(v, n) for k, v in cols.items()}
        # cal_permutation (circuit.rs:200-235)
        w = o.root_of_unity(n)
        roots = [pow(w, i, R) for i in range(n)]
        k1, k2 = (roots[0] + 1) % R, (roots[0] + 2) % R                    # find_cosets: 2 and 3
        coset = [roots, [r * k1 % R for r in roots], [r * k2 % R for r in roots]]
        sig = [list(coset[0]), list(coset[1]), list(coset[2])]
        for idx, g in enumerate(self.gates):
            for col, pos in enumerate((g.a, g.b, g.c)):
                if pos[0] not in (0, 1, 2):
                    raise ValueError("Invalid position")
                sig[col][idx] = coset[pos[0]][pos[1]]
        sigma_polys = [interpolate(s, n) for s in sig]
        return CompiledCircuit(n, polys, sigma_polys, k1, k2)


@dataclass
class CompiledCircuit:
    size: int
    g: dict          # f_a f_b f_c q_l q_r q_o q_m q_c pi as coefficient lists under a,b,c,ql,qr,qo,qm,qc,pi
    sigma: list      # s_sigma_1..3
    k1: int
    k2: int


# ----------------------------------------------------------------------------- Fiat-Shamir (challenge.rs)
def g1_serialize_uncompressed(p: Point) -> bytes:
    """ark-bls12-381 0.4 G1 `serialize_uncompressed` (zcash/IETF layout): x || y, 48-byte big-endian
    canonical integers; flag bits in byte 0 (bit 6 = infinity)."""
    if p is None:
        return bytes([0x40]) + bytes(95)
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def _rotl32(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha12_block(key_words, counter: int, double_rounds: int = 6):
    """rand_chacha ChaCha12: 64-bit block counter in words 12-13, stream id 0 in words 14-15.
    (double_rounds = 10 is ChaCha20: lets the KAT tests pin the block function on RFC 8439 vectors.)"""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
         [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(double_rounds):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


def pcg32_output(state: int) -> int:
    """PCG XSH-RR 64/32 output function (rand_core's seed expansion uses it on the advanced state)."""
    xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
    rot = state >> 59
    return ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF


def seed_from_u64(state: int) -> List[int]:
    """rand_core 0.6 `SeedableRng::seed_from_u64`: eight PCG32 steps -> the eight little-endian key words."""
    words = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & (2**64 - 1)
        words.append(pcg32_output(state))
    return words


class StdRngFromU64:
    """`StdRng::seed_from_u64` (rand_core 0.6 PCG32 seed expansion) + ChaCha12 word stream.
    `StdRngFromU64.from_seed(bytes32)` is `StdRng::from_seed`."""

    def __init__(self, state: Optional[int], key: Optional[List[int]] = None, double_rounds: int = 6):
        self.key = seed_from_u64(state) if key is None else list(key)
        self.double_rounds = double_rounds
        self.buf: List[int] = []
        self.ctr = 0

    @classmethod
    def from_seed(cls, seed: bytes, double_rounds: int = 6) -> "StdRngFromU64":
        assert len(seed) == 32
        return cls(None, key=[int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(8)],
                   double_rounds=double_rounds)

    def fill_bytes(self, n: int) -> bytes:
        assert n % 4 == 0
        return b"".join(self.next_u32().to_bytes(4, "little") for _ in range(n // 4))

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha12_block(self.key, self.ctr, self.double_rounds)
            self.ctr += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def fr_rand(self) -> int:
        """ark-ff `Fp::rand`: four u64 limbs (low first) ARE the Montgomery representation; the top
        bit is cleared; redraw while >= r.  Returns the canonical value."""
        while True:
            v = 0
            for i in range(4):
                v |= self.next_u64() << (64 * i)
            v &= (1 << 255) - 1
            if v < R:
                return v * pow(o.FR_MONT_R, -1, R) % R


class ChallengeGenerator:
    def __init__(self):
        self.data: Optional[bytes] = None
        self.generated = False

    def feed(self, p: Point):
        self.data = hashlib.sha256((self.data or b"") + g1_serialize_uncompressed(p)).digest()
        self.generated = False

    def generate_challenges(self, n: int) -> List[int]:
        if self.generated:
            raise RuntimeError("I'm hungry! Feed me something first")
        self.generated = True
        if self.data is None:
            raise RuntimeError("No data to generate seed from")
        rng = StdRngFromU64(int.from_bytes(self.data[:8], "little"))
        return [rng.fr_rand() for _ in range(n)]


# ----------------------------------------------------------------------------- prover (prover.rs)
@dataclass
class Proof:
    a: Point; b: Point; c: Point; z: Point
    t_lo: Point; t_mid: Point; t_hi: Point
    w_ev_x: Point; w_ev_wx: Point
    bar_a: int; bar_b: int; bar_c: int; bar_s_sigma_1: int; bar_s_sigma_2: int; bar_z_w: int
    u: int
    degree: int

    def commitments(self):
        return [self.a, self.b, self.c, self.z, self.t_lo, self.t_mid, self.t_hi, self.w_ev_x, self.w_ev_wx]

    def scalars(self):
        return [self.bar_a, self.bar_b, self.bar_c, self.bar_s_sigma_1, self.bar_s_sigma_2, self.bar_z_w]


def commit(poly, srs) -> Point:
    if not len(srs) > max(len(poly) - 1, 0):
        raise AssertionError("g1_points.len() > polynomial.degree()")
    return o.msm_evaluate_in_s(poly, srs)


def compute_acc(beta, gamma, cc: CompiledCircuit):
    """prover.rs:302-377, including the per-point Horner evaluations."""
    n = cc.size
    w = o.root_of_unity(n)
    roots = [pow(w, i, R) for i in range(n)]
    acc, pre = [1], 1
    for i in range(1, n):
        x = roots[i - 1]
        fa, fb, fc = p_eval(cc.g["a"], x), p_eval(cc.g["b"], x), p_eval(cc.g["c"], x)
        num = (fa + beta * x + gamma) * (fb + beta * cc.k1 * x + gamma) * (fc + beta * cc.k2 * x + gamma) % R
        den = (fa + beta * p_eval(cc.sigma[0], x) + gamma) * (fb + beta * p_eval(cc.sigma[1], x) + gamma) * \
              (fc + beta * p_eval(cc.sigma[2], x) + gamma) % R
        pre = pre * num % R * pow(den, -1, R) % R
        acc.append(pre)
    shifted = acc[1:] + acc[:1]
    return interpolate(acc, n), interpolate(shifted, n)


def l1_poly(n):
    return interpolate([1] + [0] * (n - 1), n)


def generate_proof(cc: CompiledCircuit, srs, blinding: Sequence[int], commit_fn=None) -> Proof:
    """`commit_fn(coeffs) -> point` replaces the pure-Python MSM (the C oracle's evaluate_in_s restatement is used by
    the larger test circuits; the reference's own circuits keep the all-Python path)."""
    global commit
    if commit_fn is not None:
        py_commit = commit
        commit = lambda poly, _srs: commit_fn(poly)  # noqa: E731
        try:
            return generate_proof(cc, srs, blinding)
        finally:
            commit = py_commit
    b1, b2, b3, b4, b5, b6, b7, b8, b9 = [x % R for x in blinding]
    n = cc.size
    w = o.root_of_unity(n)
    g = cc.g
    # Round 1 (prover.rs:68-92)
    ax = p_add(g["a"], mul_by_vanishing(trim([b2, b1]), n))
    bx = p_add(g["b"], mul_by_vanishing(trim([b4, b3]), n))
    cx = p_add(g["c"], mul_by_vanishing(trim([b6, b5]), n))
    a_c, b_c, c_c = commit(ax, srs), commit(bx, srs), commit(cx, srs)
    # Round 2 (:98-123)
    ch = ChallengeGenerator()
    ch.feed(a_c); ch.feed(b_c); ch.feed(c_c)
    beta, gamma = ch.generate_challenges(2)
    pre4 = mul_by_vanishing(trim([b9, b8, b7]), n)
    pre4w = mul_by_vanishing(trim([b9, b8 * w % R, b7 * w % R * w % R]), n)
    acc_x, acc_wx = compute_acc(beta, gamma, cc)
    z_x = p_add(pre4, acc_x)
    z_wx = p_add(pre4w, acc_wx)
    z_c = commit(z_x, srs)
    # Round 3 (:136-150, 381-444)
    ch.feed(z_c)
    (alpha,) = ch.generate_challenges(1)
    line1 = p_add(p_add(p_add(p_add(p_add(p_mul(p_mul(ax, bx), g["qm"]), p_mul(ax, g["ql"])), p_mul(bx, g["qr"])),
                              p_mul(cx, g["qo"])), g["pi"]), g["qc"])
    q1, r1 = divide_by_vanishing(line1, n)
    if r1:
        raise RuntimeError("No remainder 1")
    line2 = p_scale(p_mul(p_mul(p_mul(p_add(ax, trim([gamma, beta])), p_add(bx, trim([gamma, beta * cc.k1]))),
                                p_add(cx, trim([gamma, beta * cc.k2]))), z_x), alpha)
    line3 = p_scale(p_mul(p_mul(p_mul(p_add(p_add(ax, p_scale(cc.sigma[0], beta)), trim([gamma])),
                                      p_add(p_add(bx, p_scale(cc.sigma[1], beta)), trim([gamma]))),
                                p_add(p_add(cx, p_scale(cc.sigma[2], beta)), trim([gamma]))), z_wx), alpha)
    q23, r23 = divide_by_vanishing(p_sub(line2, line3), n)
    if r23:
        raise RuntimeError("No remainder here")
    zx2 = list(z_x)
    zx2[0] = (zx2[0] - 1) % R
    line4 = p_scale(p_mul(zx2, l1_poly(n)), alpha * alpha)
    q4, r4 = divide_by_vanishing(line4, n)
    if r4:
        raise RuntimeError("No remainder here")
    tx = p_add(p_add(q1, q23), q4)
    # SlicePoly::new (slice_polynomial.rs:22-43)
    tmp = len(tx) // 3
    if tmp * 3 < len(tx):
        tmp += 1
    slices = [trim(tx[i * tmp:(i + 1) * tmp]) for i in range(3)] if tmp else [[], [], []]
    degree = tmp - 1
    t_lo, t_mid, t_hi = (commit(s, srs) for s in slices)
    # Round 4 (:156-178)
    ch.feed(t_lo); ch.feed(t_mid); ch.feed(t_hi)
    (zeta,) = ch.generate_challenges(1)
    bar_a, bar_b, bar_c = p_eval(ax, zeta), p_eval(bx, zeta), p_eval(cx, zeta)
    bar_s1, bar_s2 = p_eval(cc.sigma[0], zeta), p_eval(cc.sigma[1], zeta)
    bar_z_w = p_eval(z_x, zeta * w % R)
    pi_e = p_eval(g["pi"], zeta)
    tx_compact = []
    for i, s in enumerate(slices):                                    # slice_polynomial.rs:56-69
        tx_compact = p_add(tx_compact, p_scale(s, pow(zeta, (degree + 1) * i, R)))
    # Round 5 (:183-272)
    srs0 = srs[0]
    for e in (bar_a, bar_b, bar_c, bar_s1, bar_s2, bar_z_w):
        ch.feed(o.g1_mul(srs0, e))                                   # scheme.commit_para
    (v,) = ch.generate_challenges(1)
    # compute_linearisation_polynomial (:469-568)
    l1 = p_add(p_add(p_add(p_add(p_scale(g["qm"], bar_a * bar_b), p_scale(g["ql"], bar_a)), p_scale(g["qr"], bar_b)),
                     p_scale(g["qo"], bar_c)), g["qc"])
    l1 = list(l1) if l1 else [0]
    l1[0] = (l1[0] + pi_e) % R
    l1 = trim(l1)
    s2 = (bar_a + beta * zeta + gamma) * (bar_b + beta * cc.k1 * zeta + gamma) % R * (bar_c + beta * cc.k2 * zeta + gamma) % R * alpha % R
    l2 = p_scale(z_x, s2)
    s3 = (bar_a + beta * bar_s1 + gamma) * (bar_b + beta * bar_s2 + gamma) % R * bar_z_w % R * alpha % R
    tmp2 = p_scale(cc.sigma[2], beta)
    tmp2 = list(tmp2) if tmp2 else [0]
    tmp2[0] = (tmp2[0] + bar_c + gamma) % R
    l3 = p_scale(trim(tmp2), s3)
    l1_e = p_eval(l1_poly(n), zeta)
    l4 = p_scale(zx2, l1_e * alpha % R * alpha % R)
    z_h_e = (pow(zeta, n, R) - 1) % R
    l5 = p_scale(tx_compact, z_h_e)
    r_x = p_add(p_add(p_add(p_add(l1, l2), p_neg(l3)), l4), p_neg(l5))
    bar_r = p_eval(r_x, zeta)

    def sub_para(p, k):
        q = list(p)
        q[0] = (q[0] - k) % R
        return q

    wx = p_add(p_add(p_add(p_add(p_add(trim(sub_para(r_x, bar_r)), p_scale(sub_para(ax, bar_a), v)),
                                 p_scale(sub_para(bx, bar_b), v * v)), p_scale(sub_para(cx, bar_c), v ** 3)),
                     p_scale(sub_para(cc.sigma[0], bar_s1), v ** 4)), p_scale(sub_para(cc.sigma[1], bar_s2), v ** 5))
    w_ev_x, rem = divide_linear(wx, zeta)
    if rem:
        raise RuntimeError("w_ev_x was computed incorrectly")
    w_ev_wx, rem = divide_linear(trim(sub_para(z_x, bar_z_w)), zeta * w % R)
    if rem:
        raise RuntimeError("w_ev_wx was computed incorrectly")
    wc, wwc = commit(w_ev_x, srs), commit(w_ev_wx, srs)
    ch.feed(wc); ch.feed(wwc)
    (u,) = ch.generate_challenges(1)
    return Proof(a_c, b_c, c_c, z_c, t_lo, t_mid, t_hi, wc, wwc, bar_a, bar_b, bar_c, bar_s1, bar_s2, bar_z_w, u, degree)


# ----------------------------------------------------------------------------- verifier (verifier.rs) under a known secret
def verify_with_secret(cc: CompiledCircuit, srs, secret: int, proof: Proof, commit_fn=None) -> bool:
    """`commit_fn(coeffs) -> point` replaces the big-integer MSM for the eight preprocessed commitments
    (verifier.rs:173-180) when the circuit is too large for it (the GPU tests pass the engine's commit)."""
    n = cc.size
    w = o.root_of_unity(n)
    mul, add, neg = o.g1_mul, o.g1_add, o.g1_neg
    cm = commit_fn or (lambda poly: commit(poly, srs))
    q_m, q_l, q_r, q_o, q_c = (cm(cc.g[k]) for k in ("qm", "ql", "qr", "qo", "qc"))
    s1, s2, s3 = (cm(s) for s in cc.sigma)
    # verify_challenges (:188-220)
    ch = ChallengeGenerator()
    ch.feed(proof.a); ch.feed(proof.b); ch.feed(proof.c)
    beta, gamma = ch.generate_challenges(2)
    ch.feed(proof.z)
    (alpha,) = ch.generate_challenges(1)
    ch.feed(proof.t_lo); ch.feed(proof.t_mid); ch.feed(proof.t_hi)
    (zeta,) = ch.generate_challenges(1)
    for e in proof.scalars():
        ch.feed(mul(srs[0], e))
    (v,) = ch.generate_challenges(1)
    ch.feed(proof.w_ev_x); ch.feed(proof.w_ev_wx)
    (u,) = ch.generate_challenges(1)
    if u != proof.u:
        return False
    z_h_e = (pow(zeta, n, R) - 1) % R
    l_1_e = z_h_e * pow(n * (zeta - 1) % R, -1, R) % R
    p_i_e = p_eval(cc.g["pi"], zeta)
    ba, bb, bc, bs1, bs2, bzw = proof.scalars()
    r_0 = (p_i_e - l_1_e * alpha * alpha - alpha * (ba + bs1 * beta + gamma) * (bb + bs2 * beta + gamma) * (bc + gamma) * bzw) % R
    d1 = add(add(add(add(mul(q_m, ba * bb), mul(q_l, ba)), mul(q_r, bb)), mul(q_o, bc)), q_c)
    d2 = mul(proof.z, ((ba + beta * zeta + gamma) * (bb + beta * cc.k1 * zeta + gamma) * (bc + beta * cc.k2 * zeta + gamma) * alpha
                       + l_1_e * alpha * alpha + u) % R)
    d3 = mul(s3, (ba + beta * bs1 + gamma) * (bb + beta * bs2 + gamma) * alpha * beta * bzw % R)
    d4 = mul(add(add(proof.t_lo, mul(proof.t_mid, pow(zeta, proof.degree + 1, R))),
                 mul(proof.t_hi, pow(zeta, proof.degree * 2 + 2, R))), z_h_e)
    d = add(add(d1, d2), neg(add(d3, d4)))
    f = add(add(add(add(add(d, mul(proof.a, v)), mul(proof.b, v * v)), mul(proof.c, v ** 3)), mul(s1, v ** 4)), mul(s2, v ** 5))
    e = (-r_0 + v * ba + v * v * bb + v ** 3 * bc + v ** 4 * bs1 + v ** 5 * bs2 + u * bzw) % R
    e_pt = mul(srs[0], e)
    left = mul(add(proof.w_ev_x, mul(proof.w_ev_wx, u)), secret)                      # e(., s G2)
    right = add(add(add(mul(proof.w_ev_x, zeta), mul(proof.w_ev_wx, u * zeta * w % R)), f), neg(e_pt))  # e(., G2)
    return left == right


# ----------------------------------------------------------------------------- the reference's test circuits (verifier.rs:232-382)
def circuit_accepted_01(wrong: bool = False) -> Circuit:
    c = Circuit()
    c.add_multiplication_gate((1, 0, 3), (0, 0, 3), (0, 3, 9), 0)
    c.add_multiplication_gate((1, 1, 4), (0, 1, 4), (1, 3, 16), 0)
    c.add_multiplication_gate((1, 2, 5), (0, 2, 5), (2, 3, 25), 0)
    c.add_addition_gate((2, 0, 9), (2, 1, 16), (2, 2, 20 if wrong else 25), 0)
    return c


def circuit_accepted_02() -> Circuit:
    c = Circuit()
    c.add_multiplication_gate((0, 1, 1), (1, 0, 2), (0, 3, 2), 0)
    c.add_multiplication_gate((1, 1, 1), (0, 0, 1), (0, 2, 1), 0)
    c.add_multiplication_gate((2, 1, 1), (2, 6, 3), (1, 3, 3), 0)
    c.add_addition_gate((0, 4, 2), (2, 2, 3), (0, 5, 5), 0)
    c.add_multiplication_gate((2, 0, 2), (1, 4, 3), (1, 5, 6), 0)
    c.add_addition_gate((2, 3, 5), (2, 4, 6), (2, 5, 11), 0)
    c.add_constant_gate((0, 6, 3), (1, 6, 0), (1, 2, 3), 0)
    return c


def circuit_accepted_03() -> Circuit:
    c = Circuit()
    c.add_multiplication_gate((0, 0, 1), (1, 0, 2), (0, 1, 2), 0)
    c.add_multiplication_gate((2, 0, 2), (1, 1, 3), (2, 1, 6), 0)
    return c
