"""Status codes of the C ABI on bad input (nothing throws or aborts; the Rust shim maps non-zero to panic!), and the
degenerate sizes the reference's API admits (empty polynomial, single point, size-1 domain)."""
import ctypes

import numpy as np
import pytest


def test_status_codes(zkp, engine, coracle):
    F = zkp.fields
    lib, h = engine.lib, engine._h
    out = np.zeros(12, dtype=np.uint64)
    inf = ctypes.c_uint8(0)
    s = F.random_fr_mont(1, 8)
    b = coracle.srs(F.fr_to_mont_array([3]), 8)
    engine.srs_upload(b[:4])
    # n > SRS length: scheme.rs:86 assert
    assert lib.zkp_msm_g1(h, s.ctypes.data, 8, out.ctypes.data, ctypes.byref(inf)) == 4
    # null pointers / null context
    assert lib.zkp_msm_g1(h, None, 4, out.ctypes.data, ctypes.byref(inf)) == 1
    assert lib.zkp_msm_g1(None, s.ctypes.data, 4, out.ctypes.data, ctypes.byref(inf)) == 1
    assert lib.zkp_ntt_fr(h, None, 4, 1, 0, None) == 1
    assert lib.zkp_ctx_set_msm_window(h, 1) == 1 and lib.zkp_ctx_set_msm_window(h, 23) == 1
    assert lib.zkp_sort_pairs_dev(h, None, None, 5, 8, 0) == 1
    assert lib.zkp_fr_scan_dev(h, s.ctypes.data, 8, 2, 0) == 1
    # more than 16 commitments in one batch
    ptrs = (ctypes.c_void_p * 17)(*[s.ctypes.data] * 17)
    lens = (ctypes.c_size_t * 17)(*[4] * 17)
    big = np.zeros((17, 12), dtype=np.uint64)
    assert lib.zkp_msm_g1_multi_dev(h, 17, ptrs, lens, big.ctypes.data, None) == 1
    # empty polynomial in open: scheme.rs:112 expect("at least 1")
    z = F.fr_to_mont_array([5])
    y = np.zeros(4, dtype=np.uint64)
    assert lib.zkp_kzg_open(h, s.ctypes.data, 0, z.ctypes.data, out.ctypes.data, ctypes.byref(inf), y.ctypes.data) == 7
    zero = np.zeros((3, 4), dtype=np.uint64)  # all-zero coefficients trim to the empty polynomial
    assert lib.zkp_kzg_open(h, zero.ctypes.data, 3, z.ctypes.data, out.ctypes.data, ctypes.byref(inf), y.ctypes.data) == 7
    assert lib.zkp_strerror(7).startswith(b"empty polynomial")


def test_degenerate_sizes(zkp, engine, coracle, pyref):
    F = zkp.fields
    b = coracle.srs(F.fr_to_mont_array([9]), 4)
    engine.srs_upload(b)
    # empty sum is the identity (scheme.rs:94 unwrap_or(G1Point::zero()))
    out, inf = engine.msm(np.zeros((0, 4), dtype=np.uint64))
    assert inf and not out.any()
    # one term
    s = F.fr_to_mont_array([12345])
    out, inf = engine.msm(s)
    assert F.g1_from_array(out)[0] == pyref.g1_mul(F.g1_from_array(b[0])[0], 12345)
    # size-1 domain: identity in both directions
    d = F.fr_to_mont_array([77])
    engine.ntt(d, 0)
    engine.ntt(d, 0, inverse=True)
    assert F.fr_from_mont_array(d) == [77]
    # a one-pair sort and a one-element scan
    k, v = np.array([3], dtype=np.uint32), np.array([9], dtype=np.uint32)
    if engine.lib._name.endswith("_emu.so"):
        engine.sort_pairs_dev(k, v, 1, 8)
        assert k[0] == 3 and v[0] == 9
    # constant polynomial opened anywhere: quotient zero -> identity witness
    scheme = zkp.KzgScheme(engine, zkp.Srs(b))
    op = scheme.open([42], 7)
    assert op.evaluation == 42 and op.point is None
