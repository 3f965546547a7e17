"""Device-resident Fr vector operations (csrc/poly.cu) against Python integers: the O(n) polynomial work of
the prover between its MSMs and FFTs.  Same body on the CPU kernel emulator and (gpu-marked) on the B200."""
import random

import pytest


def _rand(n, seed, R):
    rng = random.Random(seed)
    return [rng.randrange(R) for _ in range(n)]


def _vec(zkp, engine, ints):
    return engine.vec(zkp.fields.fr_to_mont_array(ints))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000])
def test_powers(zkp, engine, pyref, n):
    R = pyref.R
    base, first = 0x1234567890ABCDEF123, 77
    v = engine.vec(n=n)
    engine.fr_powers(v, base, first)
    assert v.ints() == [first * pow(base, i, R) % R for i in range(n)]


@pytest.mark.parametrize("n", [1, 15, 16, 17, 700])
def test_batch_inverse(zkp, engine, pyref, n):
    R = pyref.R
    a = [x or 1 for x in _rand(n, n, R)]
    v = _vec(zkp, engine, a)
    engine.fr_batch_inverse(v)
    assert v.ints() == [pow(x, -1, R) for x in a]


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 5000])
@pytest.mark.parametrize("op", ["mul", "add"])
@pytest.mark.parametrize("reverse", [False, True])
def test_scan(zkp, engine, pyref, n, op, reverse):
    """Multi-level inclusive scans: one block (<= 2048), block totals, two levels."""
    R = pyref.R
    a = _rand(n, 3 * n + 1, R)
    v = _vec(zkp, engine, a)
    engine.fr_scan(v, op, reverse)
    seq = a[::-1] if reverse else a
    acc, want = (1 if op == "mul" else 0), []
    for x in seq:
        acc = acc * x % R if op == "mul" else (acc + x) % R
        want.append(acc)
    assert v.ints() == (want[::-1] if reverse else want)


def test_lincomb_and_add_at(zkp, engine, pyref):
    R = pyref.R
    lens = [5, 300, 257, 0, 299]
    polys = [_rand(n, 40 + i, R) for i, n in enumerate(lens)]
    coefs = _rand(len(lens), 50, R)
    c0 = 12345
    out = engine.vec(n=310)
    vs = [_vec(zkp, engine, p) for p in polys]
    engine.fr_lincomb(out, vs, coefs, c0)
    want = [sum(c * p[j] for c, p in zip(coefs, polys) if j < len(p)) % R for j in range(310)]
    want[0] = (want[0] + c0) % R
    assert out.ints() == want
    engine.fr_add_at(out, [0, 309, 0], [5, 6, R - 1])
    want[0] = (want[0] + 5 + R - 1) % R
    want[309] = (want[309] + 6) % R
    assert out.ints() == want


def test_eval_and_trimmed_len(zkp, engine, pyref):
    R = pyref.R
    polys = [_rand(n, 60 + n, R) for n in (1, 31, 32, 33, 9000)]
    polys.append(polys[-1][:500] + [0] * 40)
    xs = _rand(len(polys), 61, R)
    vs = [_vec(zkp, engine, p) for p in polys]
    got = engine.fr_eval(vs, xs)
    assert got == [pyref.poly_eval(p, x) for p, x in zip(polys, xs)]
    assert [engine.fr_trimmed_len(v) for v in vs] == [1, 31, 32, 33, 9000, 500]
    assert engine.fr_trimmed_len(engine.vec(n=100)) == 0


def test_commit_para_batch(zkp, engine, pyref):
    """kzg/src/scheme.rs:78-82 for six scalars in one launch (the prover's round 5)."""
    srs = zkp.Srs.new_from_secret(engine, 5, 4)
    zkp.KzgScheme(engine, srs)
    ks = [0, 1, 2, pyref.R - 1] + _rand(2, 70, pyref.R)
    g0 = srs.g1_points()[0]
    assert engine.g1_mul_srs0(ks) == [pyref.g1_mul(g0, k) for k in ks]
