"""Field and curve arithmetic the kernels are built from, compiled for the host (g++), against Python
big integers.  `op 3` runs the exact even/odd carry-chain multiplication algorithm the GPU executes,
on an emulated carry flag."""
import ctypes
import random

import pytest


def _pack(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def _unpack(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_field_ops(hostlib, zkp, name):
    F = zkp.fields
    fn, mod, n, R = ((hostlib.zkp_t_fr_op, F.FR_MODULUS, 8, F.FR_R) if name == "fr"
                     else (hostlib.zkp_t_fq_op, F.FQ_MODULUS, 12, F.FQ_R))
    rinv = pow(R, -1, mod)
    rnd = random.Random(1)
    cases = [(0, 0), (0, 1), (mod - 1, mod - 1), (mod - 1, 1), (1, mod - 1), (R % mod, R % mod)]
    # half-limb patterns: the Karatsuba halves' differences change sign / vanish, carries run through whole halves
    for hi in ("ffffffff", "00000000", "80000000", "00000001", "7fffffff"):
        for lo in ("ffffffff", "00000000", "80000000", "00000001"):
            cases.append((int(hi * (n // 2) + lo * (n // 2), 16) % mod, int(lo * (n // 2) + hi * (n // 2), 16) % mod))
    cases += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(1500)]
    out = (ctypes.c_uint32 * n)()
    for a, b in cases:
        for op, exp in ((0, (a + b) % mod), (1, (a - b) % mod), (2, a * b * rinv % mod), (3, a * b * rinv % mod),
                        (7, a * b * rinv % mod), (10, a * b * rinv % mod)):  # 10 = Karatsuba + separated reduction (experiment knob)
            assert fn(op, _pack(a, n), _pack(b, n), out) == 0
            assert _unpack(out) == exp, (name, op, hex(a), hex(b))
    # dedicated squaring (triangular carry-chain rows): edge values with every top / bottom bit pattern, then random
    sq = [0, 1, 2, mod - 1, mod - 2, R % mod, (1 << 31), (1 << 32) - 1, (1 << 63) | 1, mod >> 1, (mod >> 1) + 1,
          int("55" * (4 * n), 16) % mod, int("aa" * (4 * n), 16) % mod, int("ff" * (4 * n), 16) % mod,
          int("80000000" * n, 16) % mod, int("7fffffff" * n, 16) % mod]
    sq += [rnd.randrange(mod) for _ in range(20000)]
    for a in sq if name == "fq" else []:  # Fq only: 3p < 2^384 (the rows' partial sums fit); Fr squares through fp_mul
        assert fn(8, _pack(a, n), None, out) == 0
        assert _unpack(out) == a * a * rinv % mod, (name, "sqr", hex(a))
    a = rnd.randrange(1, mod)
    fn(4, _pack(a * R % mod, n), None, out)
    assert _unpack(out) == pow(a, -1, mod) * R % mod
    fn(5, _pack(a, n), None, out)
    assert _unpack(out) == a * R % mod
    fn(6, _pack(a * R % mod, n), None, out)
    assert _unpack(out) == a


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_inverse_by_division_steps(hostlib, zkp, name):
    """csrc/inv_gcd.cuh (the inversions of the batched-affine MSM rounds, of the batch normalisations and of the prover's
    batch inverse): safegcd division steps against pow(a, -1, p) and against the Fermat ladder they replace -- small
    values, values next to p, every single-bit value (long runs of even steps), all-ones patterns, 0 -> 0, random residues."""
    F = zkp.fields
    fn, mod, n, R = ((hostlib.zkp_t_fr_op, F.FR_MODULUS, 8, F.FR_R) if name == "fr"
                     else (hostlib.zkp_t_fq_op, F.FQ_MODULUS, 12, F.FQ_R))
    bits = mod.bit_length()
    rnd = random.Random(7)
    vals = [1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, R % mod, pow(R, -1, mod), 1 << (bits - 1), (1 << (bits - 1)) - 1,
            int("55" * (4 * n), 16) % mod, int("aa" * (4 * n), 16) % mod]
    vals += [(1 << k) % mod for k in range(bits)] + [(mod - (1 << k)) % mod for k in range(bits)]
    vals += [rnd.randrange(1, mod) for _ in range(3000)]
    vals = [v for v in vals if v]
    out, out2 = (ctypes.c_uint32 * n)(), (ctypes.c_uint32 * n)()
    for a in vals:  # `a` is the stored (Montgomery) residue: a = x R, expected x^-1 R = a^-1 R^2
        assert fn(9, _pack(a, n), None, out) == 0
        assert _unpack(out) == pow(a, -1, mod) * R * R % mod, hex(a)
    for a in vals[:40]:
        fn(4, _pack(a, n), None, out2)
        fn(9, _pack(a, n), None, out)
        assert _unpack(out) == _unpack(out2)
    fn(9, _pack(0, n), None, out)
    assert _unpack(out) == 0


def test_group_law_corner_cases(hostlib, zkp, pyref):
    o = pyref
    P, R = o.P, o.FQ_MONT_R
    rnd = random.Random(2)

    def aff(p):
        if p is None:
            return _pack(0, 24)
        return _pack((p[0] * R % P) | ((p[1] * R % P) << 384), 24)

    def xyzz(p, z=1):
        if p is None:
            return _pack(0, 48)
        zz, zzz = z * z % P, z * z * z % P
        return _pack((p[0] * zz * R % P) | ((p[1] * zzz * R % P) << 384) | ((zz * R % P) << 768) | ((zzz * R % P) << 1152), 48)

    def to_aff(buf):
        out = (ctypes.c_uint32 * 24)()
        hostlib.zkp_t_g1_op(3, buf, None, out)
        v = _unpack(out)
        x, y = v & ((1 << 384) - 1), v >> 384
        if x == 0 and y == 0:
            return None
        ri = pow(R, -1, P)
        return (x * ri % P, y * ri % P)

    pts = [o.g1_mul(o.G1, rnd.randrange(o.R)) for _ in range(8)]
    out = (ctypes.c_uint32 * 48)()
    for i in range(7):
        a, b = pts[i], pts[i + 1]
        for A, B in ((a, b), (a, a), (a, o.g1_neg(a)), (None, b), (a, None), (None, None)):
            z1, z2 = rnd.randrange(1, P), rnd.randrange(1, P)
            hostlib.zkp_t_g1_op(0, xyzz(A, z1), aff(B), out)
            assert to_aff(out) == o.g1_add(A, B)
            hostlib.zkp_t_g1_op(1, xyzz(A, z1), xyzz(B, z2), out)
            assert to_aff(out) == o.g1_add(A, B)
            hostlib.zkp_t_g1_op(2, xyzz(A, z1), None, out)
            assert to_aff(out) == o.g1_add(A, A)
        k = rnd.randrange(1 << 20)
        hostlib.zkp_t_g1_op(4, xyzz(a, rnd.randrange(1, P)), _pack(k, 1), out)
        assert to_aff(out) == o.g1_mul(a, k)
