"""The reference's own kzg tests (kzg/src/commitment.rs:31-119, kzg/examples/example.rs) restated
over the drop-in mirror, run on the CPU kernel emulator and (gpu-marked) on the B200."""
import pytest

from helpers import golden, h2i, pt


def _scheme(zkp, engine, secret, size):
    srs = zkp.Srs.new_from_secret(engine, secret, size)
    return srs, zkp.KzgScheme(engine, srs)


def test_commit_known_answer(zkp, engine, pyref):
    """commitment.rs:36-54: secret = 2, size 10, poly 1 + 2X + 3X^2 -> G * p(2) = 17 G; open at 1."""
    srs, scheme = _scheme(zkp, engine, 2, 10)
    assert len(srs) == 13  # srs.rs:51: circuit_size + 3
    k = golden("kzg_msm.json")["kzg_kat"]
    assert srs.g1_points() == [pt(p) for p in k["srs"]]
    poly = [1, 2, 3]
    c = scheme.commit(poly)
    assert c.point == pt(k["commitment"])
    assert c.point == pyref.g1_mul(pyref.G1, 17)
    opening = scheme.open(poly, 1)
    assert opening.evaluation == 6
    assert opening.point == pt(k["witness"])


def test_example_rs(zkp, engine):
    """kzg/examples/example.rs with a fixed secret: x^3 + 3x + 5, opened at 4."""
    k = golden("kzg_msm.json")["kzg_example"]
    srs, scheme = _scheme(zkp, engine, h2i(k["secret"]), k["circuit_size"])
    assert srs.g1_points() == [pt(p) for p in k["srs"]]
    poly = [5, 3, 0, 1]
    assert scheme.commit(poly).point == pt(k["commitment"])
    op = scheme.open(poly, 4)
    assert op.evaluation == h2i(k["evaluation"]) == 81
    assert op.point == pt(k["witness"])


def test_scalar_mul_linearity(zkp, engine, pyref):
    """commitment.rs:61-71: commit(9 p) == 9 * commit(p)."""
    _, scheme = _scheme(zkp, engine, 0x1234567, 10)
    rng = pyref.SplitMix64(3)
    poly = [rng.fr() for _ in range(7)]
    c1 = scheme.commit(poly)
    c2 = scheme.commit([9 * x % pyref.R for x in poly])
    assert c2.point == pyref.g1_mul(c1.point, 9)


def test_aggregate_commitments(zkp, engine, pyref):
    """commitment.rs:78-89 shape: commit(p1 + v p2) == commit(p1) + v commit(p2)."""
    _, scheme = _scheme(zkp, engine, 0x777, 10)
    rng = pyref.SplitMix64(4)
    p1 = [rng.fr() for _ in range(9)]
    p2 = [rng.fr() for _ in range(9)]
    v = rng.fr()
    lhs = scheme.commit([(a + v * b) % pyref.R for a, b in zip(p1, p2)]).point
    rhs = pyref.g1_add(scheme.commit(p1).point, pyref.g1_mul(scheme.commit(p2).point, v))
    assert lhs == rhs


def test_edge_cases(zkp, engine, pyref):
    srs, scheme = _scheme(zkp, engine, 5, 4)  # 7 points
    # empty / all-zero polynomial -> identity (scheme.rs:94; nova/src/r1cs/mod.rs:52-59)
    assert scheme.commit([]).point is None
    assert scheme.commit_vector([0, 0, 0]).point is None
    # trailing zeros are trimmed before the degree assert (scheme.rs:64)
    assert scheme.commit_vector([1, 2, 0, 0, 0, 0, 0, 0, 0]).point == pyref.g1_mul(pyref.G1, 11)
    # degree >= srs length -> the reference's assert (scheme.rs:86)
    with pytest.raises(AssertionError):
        scheme.commit([1] * 8)
    # open of the empty polynomial -> expect("at least 1") (scheme.rs:112)
    with pytest.raises(ValueError):
        scheme.open([], 3)
    # constant polynomial: quotient is zero -> identity witness
    op = scheme.open([42], 9)
    assert op.evaluation == 42 and op.point is None
    # commit_para (scheme.rs:78-82)
    assert scheme.commit_para(123).point == pyref.g1_mul(pyref.G1, 123)
    # the C ABI refuses n > uploaded SRS with the dedicated status
    import numpy as np
    with pytest.raises(zkp.ZkpError) as ei:
        engine.msm(zkp.fields.fr_to_mont_array([1] * 8))
    assert ei.value.status == 4


def test_srs_generation_matches_oracle(zkp, engine, coracle):
    F = zkp.fields
    secret = 0xDEADBEEFCAFE
    srs = zkp.Srs.new_from_secret(engine, secret, 70)
    assert (srs.g1_limbs() == coracle.srs(F.fr_to_mont_array([secret]), 73)).all()


def test_open_on_device_vs_oracle(zkp, engine, coracle, pyref):
    """scheme.rs:108-120 at a size where the device path's multi-block scans / evaluation are exercised:
    (witness, y) equal the oracle's Horner + synthetic division + MSM; z = 0 and a constant polynomial too."""
    F = zkp.fields
    n = 3000
    srs = zkp.Srs.new_from_secret(engine, 0x4242, n)
    scheme = zkp.KzgScheme(engine, srs)
    rng = pyref.SplitMix64(91)
    poly = [rng.fr() for _ in range(n)] + [0, 0]  # trailing zeros are trimmed like DensePolynomial does
    pts = srs.g1_limbs()
    for z in (rng.fr(), 0, 1):
        op = scheme.open(poly, z)
        q, y = coracle.open_quotient(F.fr_to_mont_array(poly[:n]), F.fr_to_mont_array([z]))
        assert op.evaluation == F.fr_from_mont_array(y)[0] == pyref.poly_eval(poly, z)
        want = coracle.msm_pippenger(q, pts[: q.shape[0]])
        assert op.point == F.g1_from_array(want)[0]
    const = scheme.open([5], 9)
    assert const.evaluation == 5 and const.point is None  # quotient of a constant is the zero polynomial
    with pytest.raises(ValueError, match="at least 1"):
        scheme.open([0, 0], 3)
