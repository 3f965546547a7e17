"""MSM parity: golden fixtures with the corner cases, oracle on seeded inputs across window widths,
adversarial scalar distributions, generated bases."""
import numpy as np
import pytest

from helpers import affine_of, golden, h2i, pt


def test_golden_msm(zkp, engine):
    F = zkp.fields
    for case in golden("kzg_msm.json")["msm"]:
        s = F.fr_to_mont_array([h2i(x) for x in case["scalars"]])
        pts = [pt(p) for p in case["bases"]][: s.shape[0]]  # zip truncation (scheme.rs:88-91)
        b = F.g1_to_array(pts)
        out, inf = engine.msm(s, b)
        assert affine_of(zkp, out, inf) == pt(case["result"])
        # ark-ec style infinity flags instead of the (0,0) sentinel
        flags = np.array([1 if p is None else 0 for p in pts], dtype=np.uint8)
        out2, inf2 = engine.msm(s, b, infinity=flags)
        assert (out2 == out).all() and inf2 == inf


@pytest.mark.parametrize("window", [0, 3, 8, 13, 16])
def test_vs_oracle_all_windows(zkp, engine, coracle, window):
    F = zkp.fields
    n = 700
    s = F.random_fr_mont(31, n)
    b = coracle.srs(F.fr_to_mont_array([0xABCDEF]), n)
    exp = coracle.msm_pippenger(s, b)
    engine.set_msm_window(window)
    try:
        out, inf = engine.msm(s, b)
    finally:
        engine.set_msm_window(0)
    assert (out == exp).all() and not inf


def test_adversarial_scalars(zkp, engine, coracle, pyref):
    """all-zero, all-one, all r-1, all-equal scalars on one repeated point (bucket splitting and
    the P + P / P - P branches of the mixed adder)."""
    F = zkp.fields
    n = 300
    b = coracle.srs(F.fr_to_mont_array([3]), n)
    for val in (0, 1, pyref.R - 1, 0x1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F1F):
        s = F.fr_to_mont_array([val] * n)
        out, inf = engine.msm(s, b)
        assert (out == coracle.msm_pippenger(s, b)).all(), hex(val)
        assert inf == (val == 0)
    same = np.repeat(b[5:6], n, axis=0)
    s = F.fr_to_mont_array([7] * n)
    out, inf = engine.msm(s, same)
    assert affine_of(zkp, out, inf) == pyref.g1_mul(F.g1_from_array(b[5])[0], 7 * n)
    # P and -P with the same scalar cancel to the identity
    p = F.g1_from_array(b[9])[0]
    pair = F.g1_to_array([p, pyref.g1_neg(p)] * 10)
    out, inf = engine.msm(F.fr_to_mont_array([12345] * 20), pair)
    assert inf


def test_generated_bases(zkp, engine, coracle):
    """zkp_g1_generate_bases_dev: on-curve, distinct, arithmetic progression; MSM over them."""
    F = zkp.fields
    n = 200
    if engine.lib._name.endswith("_emu.so"):
        bases = np.zeros((n, 12), dtype=np.uint64)
        engine.generate_bases_dev(99, n, bases)
        sdev = None
    else:
        import torch
        t = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
        engine.generate_bases_dev(99, n, t)
        torch.cuda.synchronize()
        bases = t.cpu().numpy().view(np.uint64).reshape(n, 12)
    assert coracle.on_curve(bases)
    assert len({bytes(r) for r in bases}) == n
    s = F.random_fr_mont(41, n)
    out, inf = engine.msm(s, bases)
    assert (out == coracle.msm_pippenger(s, bases)).all()


def test_fold_partials(zkp, engine, coracle):
    """Point-range sharding on one device: two partial sums folded == the whole MSM."""
    F = zkp.fields
    n = 256
    s = F.random_fr_mont(51, n)
    b = coracle.srs(F.fr_to_mont_array([11]), n)
    whole, _ = engine.msm(s, b)
    parts = []
    for lo, hi in ((0, 100), (100, 256)):
        # host-buffer partial: stage through the resident-SRS path
        engine.srs_upload(b[lo:hi])
        o, i = engine.msm(s[lo:hi])
        # re-express the affine result as an XYZZ record with zz = zzz = 1 (Montgomery one)
        one = F.fq_to_mont_array([1])[0]
        rec = np.concatenate([o, one, one]) if not i else np.zeros(24, dtype=np.uint64)
        parts.append(rec)
    out, inf = engine.fold_partials(np.stack(parts))
    assert (out == whole).all() and not inf


@pytest.mark.parametrize("bits", [0, 3, 9, 14])
def test_fixed_base_table(zkp, engine, coracle, pyref, bits):
    """zkp_srs_precompute: every window of the signed-digit recoding reads its own 2^(c w)-shifted copy of the
    SRS and all windows share one bucket set.  Same sums as the windowed path / the oracle, for the whole SRS,
    a prefix (>= len/4 uses the table, shorter prefixes fall back to the windowed path), corner-case scalars
    and an SRS containing the point at infinity."""
    F = zkp.fields
    n = 600
    b = coracle.srs(F.fr_to_mont_array([0x5151]), n)
    b[17] = 0  # infinity base: its shifted copies must stay at infinity
    s = F.random_fr_mont(61, n)
    s[3] = 0
    s[4] = F.fr_to_mont_array([1])[0]
    s[5] = F.fr_to_mont_array([pyref.R - 1])[0]
    engine.srs_upload(b)
    engine.srs_precompute(bits)
    try:
        for m in (n, 599, 300, 150, 149, 7, 1):
            out, inf = engine.msm(s[:m])
            assert (out == coracle.msm_pippenger(s[:m], b[:m])).all() and not inf, m
        zero = np.zeros((n, 4), dtype=np.uint64)
        assert engine.msm(zero)[1]
        rm1 = F.fr_to_mont_array([pyref.R - 1] * n)
        assert (engine.msm(rm1)[0] == coracle.msm_pippenger(rm1, b)).all()
    finally:
        engine.srs_upload(b[:1])  # drops the table


def test_precompute_then_new_srs_drops_table(zkp, engine, coracle):
    F = zkp.fields
    b1 = coracle.srs(F.fr_to_mont_array([21]), 64)
    b2 = coracle.srs(F.fr_to_mont_array([22]), 64)
    s = F.random_fr_mont(71, 64)
    engine.srs_upload(b1)
    engine.srs_precompute(5)
    engine.srs_upload(b2)  # a stale table would give sums over b1
    assert (engine.msm(s)[0] == coracle.msm_pippenger(s, b2)).all()


def test_multi_msm_one_pipeline(zkp, engine, coracle):
    """zkp_msm_g1_multi_dev: several commitments against the resident SRS as bucket sets of one pipeline
    (with the fixed-base table) or one after the other (without) -- the same points either way, ragged lengths,
    an all-zero vector and an empty one included."""
    F = zkp.fields
    n = 400
    b = coracle.srs(F.fr_to_mont_array([0x7777]), n)
    vecs = [F.random_fr_mont(81, n), F.random_fr_mont(82, 399), F.random_fr_mont(83, 250),
            np.zeros((n, 4), dtype=np.uint64), F.random_fr_mont(84, 120)]
    want = [coracle.msm_pippenger(v, b[:v.shape[0]]) for v in vecs]
    engine.srs_upload(b)
    for table in (False, True):
        if table:
            engine.srs_precompute(7)
        dv = [engine.vec(v) for v in vecs]
        got = engine.msm_multi_dev([d.ptr for d in dv], [v.shape[0] for v in vecs])
        for j, (out, inf) in enumerate(got):
            if j == 3:
                assert inf
            else:
                assert (out == want[j]).all() and not inf, (table, j)
        # with an empty member the batch falls back to one MSM at a time; the empty sum is the identity
        got = engine.msm_multi_dev([dv[0].ptr, dv[0].ptr], [n, 0])
        assert (got[0][0] == want[0]).all() and got[1][1]
    engine.srs_upload(b[:1])


@pytest.mark.parametrize("rounds", [1, 2, 3, 5, 12])
def test_batched_affine_rounds(zkp, engine, coracle, pyref, rounds):
    """csrc/msm_affine.cu: R batched-affine tree rounds before the XYZZ finish return the same point as R = 0 and the
    oracle -- random scalars at two window widths (long and short runs, odd run lengths, empty buckets), the fixed-base
    table, and every special case of the affine adder: infinity bases, one point repeated (P + P in every round until
    the run is gone), P and -P (cancellation to O inside a run), all-equal scalars (one bucket holds everything; more
    rounds than the run is deep)."""
    F = zkp.fields
    n = 500
    b = coracle.srs(F.fr_to_mont_array([0xAFF1]), n)
    b[11] = 0
    b[12] = 0
    s = F.random_fr_mont(91, n)
    s[7] = 0
    engine.set_msm_affine(rounds)
    try:
        for window in (4, 9):
            engine.set_msm_window(window)
            out, inf = engine.msm(s, b)
            assert engine.last_affine_rounds() == rounds
            assert (out == coracle.msm_pippenger(s, b)).all() and not inf, window
        engine.set_msm_window(0)
        # one bucket per window holds every point
        for val in (1, 5, pyref.R - 1):
            sv = F.fr_to_mont_array([val] * n)
            assert (engine.msm(sv, b)[0] == coracle.msm_pippenger(sv, b)).all(), val
        if rounds == 2:  # one bucket owns more than AFF_TB_SERIAL threads of the round: the block-filled start-bucket list
            big = np.tile(b[20:120], (26, 1))
            sv = F.fr_to_mont_array([3] * big.shape[0])
            assert (engine.msm(sv, big)[0] == coracle.msm_pippenger(sv, big)).all()
        # the same point n times: doublings all the way down
        same = np.repeat(b[5:6], 333, axis=0)
        out, inf = engine.msm(F.fr_to_mont_array([9] * 333), same)
        assert affine_of(zkp, out, inf) == pyref.g1_mul(F.g1_from_array(b[5])[0], 9 * 333)
        # P, -P, P, -P ...: every pair cancels; then with one extra P
        p = F.g1_from_array(b[9])[0]
        pair = F.g1_to_array([p, pyref.g1_neg(p)] * 10)
        assert engine.msm(F.fr_to_mont_array([12345] * 20), pair)[1]
        odd = F.g1_to_array([p, pyref.g1_neg(p)] * 10 + [p])
        out, inf = engine.msm(F.fr_to_mont_array([12345] * 21), odd)
        assert affine_of(zkp, out, inf) == pyref.g1_mul(p, 12345)
        # fixed-base table + a batch of commitments as bucket sets
        engine.srs_upload(b)
        engine.srs_precompute(6)
        assert (engine.msm(s)[0] == coracle.msm_pippenger(s, b)).all()
        vecs = [F.random_fr_mont(92, n), F.random_fr_mont(93, 321)]
        dv = [engine.vec(v) for v in vecs]
        got = engine.msm_multi_dev([d.ptr for d in dv], [v.shape[0] for v in vecs])
        for j, (o, i) in enumerate(got):
            assert (o == coracle.msm_pippenger(vecs[j], b[:vecs[j].shape[0]])).all() and not i, j
    finally:
        engine.set_msm_affine(-1)
        engine.set_msm_window(0)
        engine.srs_upload(b[:1])
