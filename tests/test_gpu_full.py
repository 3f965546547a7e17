"""GPU-only parity at BASELINE.json sizes: full compares where the oracle finishes in seconds, and
size-independent properties (round trips, shard-fold consistency, closed-form sums) at 2^24."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def _host(t, cols):
    return t.cpu().numpy().view(np.uint64).reshape(-1, cols)


@pytest.mark.parametrize("log_n", [16, 19, 20, 22])
def test_ntt_large_vs_oracle(zkp, gpu_engine, coracle, log_n):
    """three-pass schedules (>= 2^19) element-for-element against the C oracle."""
    F = zkp.fields
    a = F.random_fr_mont(0x600 + log_n, 1 << log_n)
    t = _dev(a)
    gpu_engine.ntt_dev(t, log_n)
    assert (_host(t, 4) == coracle.ntt(a, log_n)).all()
    gpu_engine.ntt_dev(t, log_n, inverse=True)
    assert (_host(t, 4) == a).all()
    gpu_engine.ntt_dev(t, log_n, coset=7)
    assert (_host(t, 4) == coracle.ntt(a, log_n, False, F.fr_to_mont_array([7]))).all()


def test_ntt_2p24_properties(zkp, gpu_engine, coracle):
    """2^24 (the metric's size): full compare against the multi-threaded oracle + round trip + batch."""
    import torch

    F = zkp.fields
    log_n = 24
    a = F.random_fr_mont(0x2424, 1 << log_n)
    t = _dev(a)
    gpu_engine.ntt_dev(t, log_n)
    got = _host(t, 4)
    assert (got == coracle.ntt(a, log_n)).all()
    gpu_engine.ntt_dev(t, log_n, inverse=True)
    assert (_host(t, 4) == a).all()
    # batch of 4 x 2^22 equals four single transforms
    tb = _dev(a)
    gpu_engine.ntt_dev(tb, 22, batch=4)
    one = _dev(a[(1 << 22):(2 << 22)])
    gpu_engine.ntt_dev(one, 22)
    assert torch.equal(tb.view(4, -1)[1], one.reshape(-1))


@pytest.mark.parametrize("log_n", [14, 18, 20, 22])
def test_msm_large_vs_oracle(zkp, gpu_engine, coracle, log_n):
    import torch

    F = zkp.fields
    n = 1 << log_n
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(0x5000 + log_n, n, bases)
    s = F.random_fr_mont(0x7000 + log_n, n)
    out, inf = gpu_engine.msm_dev(_dev(s), bases, n)
    exp = coracle.msm_pippenger(s, _host(bases, 12))
    assert (out == exp).all() and not inf


def _splitmix(seed):
    s = seed & (2**64 - 1)
    while True:
        s = (s + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        yield z ^ (z >> 31)


def _progression(seed):
    """(a0, delta) of zkp_g1_generate_bases_dev: point i = (a0 + i * delta) * G."""
    g = _splitmix(seed)
    a0 = dl = 0
    for i in range(4):
        x, y = next(g), next(g)
        a0 |= x << (64 * i)
        dl |= y << (64 * i)
    mask = (1 << 254) - 1
    return a0 & mask, (dl & mask) | 1


def test_msm_2p24_closed_form_and_shards(zkp, gpu_engine, pyref):
    """2^24 points (the metric's size).  With every scalar equal to k the sum over the generated
    progression has a closed form, k * (n a0 + delta n(n-1)/2) * G, computed on the CPU in
    microseconds; then random scalars: whole MSM == fold of two point-range shards."""
    import torch

    F = zkp.fields
    n = 1 << 24
    seed = 0x2424
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(seed, n, bases)
    a0, dl = _progression(seed)
    k = 0x1234567
    s = _dev(np.repeat(F.fr_to_mont_array([k]), n, axis=0))
    out, inf = gpu_engine.msm_dev(s, bases, n)
    exp = pyref.g1_mul(pyref.G1, k * (n * a0 + dl * (n * (n - 1) // 2)) % pyref.R)
    assert F.g1_from_array(out)[0] == exp and not inf
    del s
    # random scalars: shard consistency
    sr = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
    whole, _ = gpu_engine.msm_dev(sr, bases, n)
    h = n // 2 + 12345
    p0 = gpu_engine.msm_partial_dev(sr, bases, h)
    p1 = gpu_engine.msm_partial_dev(sr[h * 4:], bases[h * 12:], n - h)
    folded, _ = gpu_engine.fold_partials(np.stack([p0, p1]))
    assert (folded == whole).all()


def test_e2e_host_buffers(zkp, gpu_engine, coracle):
    """The reference-facing entry points with host buffers (what the Rust shim calls)."""
    F = zkp.fields
    n = 1 << 16
    s = F.random_fr_mont(0x16, n)
    srs = zkp.Srs.new_from_secret(gpu_engine, 0xFACE, n - 3)  # n points
    scheme = zkp.KzgScheme(gpu_engine, srs)
    out, inf = gpu_engine.msm(s)
    assert (out == coracle.msm_pippenger(s, srs.g1_limbs())).all()
    a = F.random_fr_mont(0x17, 1 << 18)
    d = a.copy()
    gpu_engine.ntt(d, 18)
    assert (d == coracle.ntt(a, 18)).all()


@pytest.mark.parametrize("log_n", [16, 20])
def test_msm_fixed_base_vs_oracle(zkp, gpu_engine, coracle, log_n):
    """The path every kzg commit takes: resident SRS + window table (zkp_srs_precompute), full compare with the
    oracle, and a batch of three commitments as one pipeline (zkp_msm_g1_multi_dev)."""
    import torch

    F = zkp.fields
    n = 1 << log_n
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(0x9000 + log_n, n, bases)
    gpu_engine.srs_upload_dev(bases, n)
    gpu_engine.srs_precompute()
    hb = _host(bases, 12)
    vecs = [F.random_fr_mont(0xA000 + log_n + j, n - 3 * j) for j in range(3)]
    out, inf = gpu_engine.msm_dev(_dev(vecs[0]), None, n)
    want0 = coracle.msm_pippenger(vecs[0], hb)
    assert (out == want0).all() and not inf
    dv = [_dev(v) for v in vecs]
    got = gpu_engine.msm_multi_dev(dv, [v.shape[0] for v in vecs])
    for j, (o, i) in enumerate(got):
        assert (o == coracle.msm_pippenger(vecs[j], hb[:vecs[j].shape[0]])).all() and not i, j
    gpu_engine.srs_upload_dev(bases, 1)


@pytest.mark.parametrize("window", [13, 15, 16, 17, 20])
def test_msm_fixed_base_window_widths(zkp, gpu_engine, coracle, window):
    """Fixed-base tables of several window widths at 2^17 points against the oracle: widths whose top window holds only a
    few bits (15: the carry alone) pile every point into a handful of buckets (heavily split runs, the block-filled
    start-bucket list of the affine rounds), 17 and 20 take three sort passes and a deeper reduction tree, and forced
    affine rounds run on the short runs of the wide windows."""
    import torch

    F = zkp.fields
    n = 1 << 17
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(0x9100, n, bases)
    gpu_engine.srs_upload_dev(bases, n)
    gpu_engine.srs_precompute(window)
    s = F.random_fr_mont(0xA100 + window, n)
    want = coracle.msm_pippenger(s, _host(bases, 12))
    try:
        for rounds in (-1, 2):
            gpu_engine.set_msm_affine(rounds)
            out, inf = gpu_engine.msm_dev(_dev(s), None, n)
            assert gpu_engine.last_msm_shape()[0] == window
            assert (out == want).all() and not inf, (window, rounds)
    finally:
        gpu_engine.set_msm_affine(-1)
        gpu_engine.srs_upload_dev(bases, 1)


def test_msm_2p24_fixed_base_properties(zkp, gpu_engine, pyref):
    """2^24 points through the fixed-base table (the bench's step): closed-form sum with equal scalars, and random
    scalars equal to the windowed path over the same points (two different bucket layouts, same group element)."""
    import torch

    F = zkp.fields
    n = 1 << 24
    seed = 0x2425
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(seed, n, bases)
    gpu_engine.srs_upload_dev(bases, n)
    gpu_engine.srs_precompute()
    a0, dl = _progression(seed)
    k = pyref.R - 0x1234567  # large scalar: every window non-trivial, negative digits
    s = _dev(np.repeat(F.fr_to_mont_array([k]), n, axis=0))
    out, inf = gpu_engine.msm_dev(s, None, n)
    exp = pyref.g1_mul(pyref.G1, k * (n * a0 + dl * (n * (n - 1) // 2)) % pyref.R)
    assert F.g1_from_array(out)[0] == exp and not inf
    del s
    sr = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
    fixed, _ = gpu_engine.msm_dev(sr, None, n)
    assert gpu_engine.last_msm_shape()[0] >= 20
    windowed, _ = gpu_engine.msm_dev(sr, bases, n)
    assert (fixed == windowed).all()
    gpu_engine.srs_upload_dev(bases, 1)  # drop the 18 GiB table


def test_msm_2p24_random_scalars_vs_oracle(zkp, gpu_engine, coracle):
    """The metric's own size, random scalars, FULL compare: 2^24-point MSM through the fixed-base table (the bench's
    step, `KzgScheme::commit`'s path) and through the windowed path (ad-hoc bases) against the multi-threaded C oracle
    (`orc_msm_pippenger`, ~15 s on 16 cores) -- kzg/src/scheme.rs:84-96 on the same buffers."""
    import torch

    F = zkp.fields
    n = 1 << 24
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    gpu_engine.generate_bases_dev(0x2426, n, bases)
    s = F.random_fr_mont(0x2427, n)
    want = coracle.msm_pippenger(s, _host(bases, 12))
    sd = _dev(s)
    windowed, inf = gpu_engine.msm_dev(sd, bases, n)
    assert (windowed == want).all() and not inf
    gpu_engine.srs_upload_dev(bases, n)
    gpu_engine.srs_precompute()
    fixed, inf = gpu_engine.msm_dev(sd, None, n)
    assert gpu_engine.last_msm_shape()[0] >= 20
    assert (fixed == want).all() and not inf
    gpu_engine.srs_upload_dev(bases, 1)  # drop the table


def test_affine_rounds_fall_back_when_memory_is_short(zkp, coracle):
    """The round buffers of the batched-affine accumulation are all-or-nothing: with too little free HBM the MSM runs
    XYZZ-only (rounds = 0) and returns the same point.  A fresh context (its scratch buffers start empty)."""
    import torch

    F = zkp.fields
    eng = zkp.Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 1 << 22
    bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
    eng.generate_bases_dev(0x5153, n, bases)
    s = F.random_fr_mont(0x5154, n)
    sd = _dev(s)
    want = coracle.msm_pippenger(s, _host(bases, 12))
    eng.set_msm_affine(3)
    hog = None
    try:
        torch.cuda.empty_cache()
        free, _total = torch.cuda.mem_get_info()
        hog = torch.empty(max(free - (5 << 30), 0), dtype=torch.uint8, device="cuda")  # less than the guard's 6 GiB margin
        out, _ = eng.msm_dev(sd, bases, n)
        assert eng.last_affine_rounds() == 0 and (out == want).all()
        del hog
        hog = None
        torch.cuda.empty_cache()
        out, _ = eng.msm_dev(sd, bases, n)
        assert eng.last_affine_rounds() == 3 and (out == want).all()
    finally:
        del hog
        torch.cuda.empty_cache()
        eng.close()
