"""NTT / iNTT / coset parity: golden fixtures, C oracle on seeded inputs, padding, batches, products."""
import numpy as np
import pytest

from helpers import golden, h2i


def test_golden_vectors(zkp, engine):
    F = zkp.fields
    for case in golden("ntt.json")["ntt"]:
        n = case["log_n"]
        src = F.fr_to_mont_array([h2i(x) for x in case["input"]])
        cs = h2i(case["coset"])
        for inverse, coset, key in ((False, None, "fft"), (True, None, "ifft"), (False, cs, "coset_fft"), (True, cs, "coset_ifft")):
            d = src.copy()
            engine.ntt(d, n, 1, inverse=inverse, coset=coset)
            assert F.fr_from_mont_array(d) == [h2i(x) for x in case[key]], (n, key)


@pytest.mark.parametrize("log_n", [1, 4, 7, 10, 11, 12, 13])
def test_vs_oracle_seeded(zkp, engine, coracle, log_n):
    """single-tile (<= 2^11) and two-pass schedules, all four transform kinds."""
    F = zkp.fields
    a = F.random_fr_mont(0xB200 + log_n, 1 << log_n)
    h = F.fr_to_mont_array([7])
    for inverse in (False, True):
        for coset in (None, 7):
            d = a.copy()
            engine.ntt(d, log_n, 1, inverse=inverse, coset=coset)
            assert (d == coracle.ntt(a, log_n, inverse, None if coset is None else h)).all(), (log_n, inverse, coset)


def test_round_trip_and_coset_one(zkp, engine):
    F = zkp.fields
    a = F.random_fr_mont(9, 1 << 12)
    d = a.copy()
    engine.ntt(d, 12)
    engine.ntt(d, 12, inverse=True)
    assert (d == a).all()
    # offset 1 is the plain domain
    d1, d2 = a.copy(), a.copy()
    engine.ntt(d1, 12, coset=1)
    engine.ntt(d2, 12)
    assert (d1 == d2).all()


def test_batched(zkp, engine, coracle):
    F = zkp.fields
    batch, log_n = 5, 12
    a = F.random_fr_mont(11, batch << log_n)
    d = a.copy()
    engine.ntt(d, log_n, batch)
    for i in range(batch):
        assert (d.reshape(batch, -1, 4)[i] == coracle.ntt(a.reshape(batch, -1, 4)[i], log_n)).all()


def test_padded_interpolate(zkp, engine):
    """`Evaluations::interpolate` on a vector shorter than the domain (plonk/src/circuit.rs:131-133):
    the caller zero-pads, as ark-poly's `ifft_in_place` resize does."""
    F = zkp.fields
    case = golden("ntt.json")["padded_interpolate"]
    v = [h2i(x) for x in case["input"]]
    d = F.fr_to_mont_array(v + [0] * ((1 << case["log_n"]) - len(v)))
    engine.ntt(d, case["log_n"], inverse=True)
    assert F.fr_from_mont_array(d) == [h2i(x) for x in case["ifft"]]


def test_poly_mul(zkp, engine, coracle):
    """`&DensePolynomial * &DensePolynomial` (plonk/src/prover.rs:396-437): golden + oracle + zero cases."""
    F = zkp.fields
    pm = golden("ntt.json")["poly_mul"]
    out = engine.poly_mul(F.fr_to_mont_array([h2i(x) for x in pm["a"]]), F.fr_to_mont_array([h2i(x) for x in pm["b"]]))
    assert F.fr_from_mont_array(out) == [h2i(x) for x in pm["product"]]
    a, b = F.random_fr_mont(21, 1500), F.random_fr_mont(22, 3000)  # product domain 2^13: two-pass NTT
    assert (engine.poly_mul(a, b) == coracle.poly_mul(a, b)).all()
    assert engine.poly_mul(a[:0], b).shape[0] == 0
    one = F.fr_to_mont_array([1])
    assert (engine.poly_mul(a, one) == a).all()


def test_domain_too_large(zkp, engine):
    """ark-poly `GeneralEvaluationDomain::new` returns None past 2^32; this engine caps at 2^27."""
    import ctypes

    buf = np.zeros(4, dtype=np.uint64)
    st = engine.lib.zkp_ntt_fr(engine._h, ctypes.c_void_p(buf.ctypes.data), 33, 1, 0, None)
    assert st == 5
    with pytest.raises(zkp.ZkpError):
        engine._check(st)
