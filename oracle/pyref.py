"""Pure-Python big-integer restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

This file is the *independent cross-oracle*: textbook affine curve formulas and Python ``int``
arithmetic, no Montgomery limbs, no shared code with the C oracle (``oracle/zkp_oracle.c``) or the
CUDA engine.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import it; the product path never does.

What it restates (reference file:line, relative to sota-zk-labs/zkp-implementation):
  * ``KzgScheme::evaluate_in_s``   kzg/src/scheme.rs:84-96   -> :func:`msm_evaluate_in_s`
  * ``KzgScheme::open``            kzg/src/scheme.rs:108-120 -> :func:`kzg_open`
  * ``Srs::new_from_secret``       kzg/src/srs.rs:48-69      -> :func:`srs_from_secret`
  * ark-poly ``Radix2EvaluationDomain`` fft/ifft/coset (reached from plonk/src/prover.rs:374-375,
    396-437 and plonk/src/circuit.rs:175,230-232)            -> :func:`ntt`, :func:`intt`

The arithmetic itself lives in un-vendored third-party crates (ark-ff/ark-ec/ark-poly 0.4.2,
ark-bls12-381 0.4.0; no Cargo.lock in the reference), so their *published* semantics are restated:
natural-order in/out NTT with omega_N = 7^((r-1)/N), 1/N on the inverse, coset offset h applied as
x[j]*h^j before the forward transform and h^-j after the inverse one; MSM result compared as the
unique normalised affine point.

Parity status: the reference holds NO literal golden vectors.  The only
deterministic known-answer test is kzg/src/commitment.rs:36-54 (secret=2, commit(1+2X+3X^2)=17*G),
which :func:`selfcheck` reproduces, together with the public BLS12-381 constants.  Absolute-value
parity against a Rust run is therefore "pinned by algebra + public constants only".
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- constants
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1 = (GX, GY)
INF = None  # point at infinity
FR_GENERATOR = 7
TWO_ADICITY = 32
ROOT_2_32 = pow(FR_GENERATOR, (R - 1) >> TWO_ADICITY, R)  # ark-ff TWO_ADIC_ROOT_OF_UNITY
FR_MONT_R = (1 << 256) % R
FQ_MONT_R = (1 << 384) % P
MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------- seeded PRNG (splitmix64)
class SplitMix64:
    """Deterministic 64-bit generator used for every synthetic workload in this repo."""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def fr(self) -> int:
        """Uniform canonical Fr element: 4 limbs (low first), top bit cleared, rejection-sampled."""
        while True:
            v = 0
            for i in range(4):
                v |= self.next() << (64 * i)
            v &= (1 << 255) - 1
            if v < R:
                return v


# ----------------------------------------------------------------------------- G1 (affine, textbook)
def is_on_curve(pt) -> bool:
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def g1_neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    """Affine + affine with all corner cases (ark-ec ``Affine + Affine``)."""
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt, k: int):
    """MSB-first double-and-add over the canonical integer (ark-ec ``mul_bigint``)."""
    k %= R
    acc = INF
    for bit in bin(k)[2:] if k else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, pt)
    return acc


# ----------------------------------------------------------------------------- KZG (kzg/src)
def srs_from_secret(secret: int, circuit_size: int):
    """kzg/src/srs.rs:48-69 -- circuit_size + 3 powers of the secret times G (affine)."""
    out, cur = [], 1
    for _ in range(circuit_size + 3):
        out.append(g1_mul(G1, cur))
        cur = cur * secret % R
    return out


def msm_evaluate_in_s(coeffs, points):
    """kzg/src/scheme.rs:84-96 -- zip-truncated sum of per-term scalar muls; empty -> identity."""
    acc = INF
    for c, pt in zip(coeffs, points):
        acc = g1_add(acc, g1_mul(pt, c))
    return acc


def poly_trim(coeffs):
    """DensePolynomial::from_coefficients_vec truncates trailing zeros."""
    c = list(coeffs)
    while c and c[-1] % R == 0:
        c.pop()
    return c


def poly_eval(coeffs, z: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R
    return acc


def kzg_open_quotient(coeffs, z: int):
    """kzg/src/scheme.rs:108-117 -- y = p(z); q = (p - y) / (X - z) by synthetic division."""
    if not coeffs:
        raise ValueError("at least 1")  # scheme.rs:112 expect("at least 1")
    y = poly_eval(coeffs, z)
    n = len(coeffs)
    q = [0] * (n - 1)
    carry = 0
    for i in range(n - 1, 0, -1):
        carry = (coeffs[i] + carry * z) % R
        q[i - 1] = carry
    return poly_trim(q), y


def kzg_open(coeffs, z: int, points):
    q, y = kzg_open_quotient(coeffs, z)
    return msm_evaluate_in_s(q, points), y


# ----------------------------------------------------------------------------- NTT (ark-poly Radix2)
def root_of_unity(n: int) -> int:
    """group_gen of GeneralEvaluationDomain::<Fr>::new(n) for a power of two n."""
    log_n = n.bit_length() - 1
    assert 1 << log_n == n and log_n <= TWO_ADICITY
    return pow(ROOT_2_32, 1 << (TWO_ADICITY - log_n), R)


def _fft_rec(a, w):
    n = len(a)
    if n == 1:
        return a
    e = _fft_rec(a[0::2], w * w % R)
    o = _fft_rec(a[1::2], w * w % R)
    out = [0] * n
    t = 1
    h = n // 2
    for k in range(h):
        v = t * o[k] % R
        out[k] = (e[k] + v) % R
        out[k + h] = (e[k] - v) % R
        t = t * w % R
    return out


def ntt(values, log_n: int, coset: int | None = None):
    """fft_in_place: zero-pad to 2^log_n, out[i] = sum_j v[j] (h w^i)^j, natural order."""
    n = 1 << log_n
    a = [v % R for v in values] + [0] * (n - len(values))
    assert len(a) == n
    if coset is not None:
        hp = 1
        for j in range(n):
            a[j] = a[j] * hp % R
            hp = hp * coset % R
    return _fft_rec(a, root_of_unity(n))


def intt(values, log_n: int, coset: int | None = None):
    """ifft_in_place: out[j] = N^-1 sum_i v[i] w^(-ij); with a coset offset, then out[j] *= h^-j."""
    n = 1 << log_n
    a = [v % R for v in values] + [0] * (n - len(values))
    assert len(a) == n
    out = _fft_rec(a, pow(root_of_unity(n), -1, R))
    ninv = pow(n, -1, R)
    out = [x * ninv % R for x in out]
    if coset is not None:
        hinv = pow(coset, -1, R)
        hp = 1
        for j in range(n):
            out[j] = out[j] * hp % R
            hp = hp * hinv % R
    return out


def ntt_naive(values, log_n: int):
    n = 1 << log_n
    w = root_of_unity(n)
    a = list(values) + [0] * (n - len(values))
    return [sum(a[j] * pow(w, i * j, R) for j in range(n)) % R for i in range(n)]


def poly_mul(a, b):
    """ark-poly ``&DensePolynomial * &DensePolynomial``: exact product, trimmed."""
    a, b = poly_trim(a), poly_trim(b)
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % R
    return poly_trim(out)


# ----------------------------------------------------------------------------- limb packing helpers
def fr_to_mont_limbs(x: int):
    v = x * FR_MONT_R % R
    return [(v >> (64 * i)) & MASK64 for i in range(4)]


def fr_from_mont_limbs(l) -> int:
    v = sum(int(l[i]) << (64 * i) for i in range(4))
    return v * pow(FR_MONT_R, -1, R) % R


def fq_to_mont_limbs(x: int):
    v = x * FQ_MONT_R % P
    return [(v >> (64 * i)) & MASK64 for i in range(6)]


def fq_from_mont_limbs(l) -> int:
    v = sum(int(l[i]) << (64 * i) for i in range(6))
    return v * pow(FQ_MONT_R, -1, P) % P


# ----------------------------------------------------------------------------- self-validation
def selfcheck() -> None:
    assert is_on_curve(G1)
    assert g1_mul(G1, R - 1) == g1_neg(G1) and g1_add(g1_mul(G1, R - 1), G1) is INF
    assert pow(ROOT_2_32, 1 << 31, R) == R - 1
    assert ROOT_2_32 == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    # kzg/src/commitment.rs:36-54: secret = 2, commit(1 + 2X + 3X^2) == 17 * G
    srs = srs_from_secret(2, 10)
    assert len(srs) == 13
    c = msm_evaluate_in_s([1, 2, 3], srs)
    assert c == g1_mul(G1, 17)
    assert c == (
        0x1098F178F84FC753A76BB63709E9BE91EEC3FF5F7F3A5F4836F34FE8A1A6D6C5578D8FD820573CEF3A01E2BFEF3EAF3A,
        0x0EA923110B733B531006075F796CC9368F2477FE26020F465468EFBB380CE1F8EEBAF5C770F31D320F9BD378DC758436,
    )
    # open at 1: q = (p - 6)/(X - 1) = 3X + 5 -> q(2) = 11
    w, y = kzg_open([1, 2, 3], 1, srs)
    assert y == 6 and w == g1_mul(G1, 11)
    # NTT vs naive DFT, inverse round trip, coset definition
    rng = SplitMix64(1)
    v = [rng.fr() for _ in range(16)]
    f = ntt(v, 4)
    assert f == ntt_naive(v, 4)
    assert intt(f, 4) == v
    h = 7
    w16 = root_of_unity(16)
    assert ntt(v, 4, coset=h) == [poly_eval(v, h * pow(w16, i, R) % R) for i in range(16)]
    assert intt(ntt(v, 4, coset=h), 4, coset=h) == v


if __name__ == "__main__":
    selfcheck()
    print("pyref selfcheck OK")
