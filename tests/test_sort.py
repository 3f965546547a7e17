"""The MSM's grouping step on its own (csrc/sort.cu): stable LSD radix sort of (key, value) pairs and the exclusive
u32 scan, against numpy.  On the emulator at small sizes (every tile / level boundary), on the GPU at MSM sizes."""
import numpy as np
import pytest


def _run_sort(engine, keys, vals, key_bits, descending):
    if engine.lib._name.endswith("_emu.so"):
        k, v = keys.copy(), vals.copy()
        engine.sort_pairs_dev(k, v, len(k), key_bits, descending)
        return k, v
    import torch
    k = torch.from_numpy(keys.view(np.int32).copy()).cuda()
    v = torch.from_numpy(vals.view(np.int32).copy()).cuda()
    engine.sort_pairs_dev(k, v, len(keys), key_bits, descending)
    torch.cuda.synchronize()
    return k.cpu().numpy().view(np.uint32), v.cpu().numpy().view(np.uint32)


def _want(keys, vals, key_bits, descending):
    masked = keys & np.uint32((1 << key_bits) - 1 if key_bits < 32 else 0xFFFFFFFF)
    sk = (~masked if descending else masked).astype(np.uint32) & np.uint32((1 << key_bits) - 1 if key_bits < 32 else 0xFFFFFFFF)
    order = np.argsort(sk, kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n,key_bits,descending", [(1, 8, False), (1000, 5, False), (1024, 8, True), (1025, 13, False),
                                                   (5000, 21, False), (3333, 11, True), (4097, 32, False)])
def test_radix_sort_pairs(zkp, engine, n, key_bits, descending):
    rng = np.random.default_rng(n + key_bits)
    hi = (1 << min(key_bits, 31)) if key_bits < 32 else (1 << 32)
    keys = rng.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 100:
        keys[10:60] = keys[5]  # long runs of equal keys: stability
    vals = np.arange(n, dtype=np.uint32)[::-1].copy()
    k, v = _run_sort(engine, keys, vals, key_bits, descending)
    wk, wv = _want(keys, vals, key_bits, descending)
    assert (k == wk).all() and (v == wv).all()


@pytest.mark.parametrize("n", [1, 1023, 1024, 1025, 70000])
def test_scan_exclusive(zkp, engine, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 1000, size=n, dtype=np.uint64).astype(np.uint32)
    want = np.concatenate([[0], np.cumsum(a[:-1], dtype=np.uint64)]).astype(np.uint32)
    if engine.lib._name.endswith("_emu.so"):
        out = np.zeros(n, dtype=np.uint32)
        engine.scan_exclusive_u32_dev(a, out, n)
    else:
        import torch
        t = torch.from_numpy(a.view(np.int32).copy()).cuda()
        o = torch.zeros_like(t)
        engine.scan_exclusive_u32_dev(t, o, n)
        torch.cuda.synchronize()
        out = o.cpu().numpy().view(np.uint32)
    assert (out == want).all()


@pytest.mark.gpu
def test_radix_sort_msm_size(zkp, gpu_engine):
    """2^26 pairs with 22-bit keys (a 2^22-point MSM's worth) against torch's stable sort."""
    import torch

    n = 1 << 26
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    keys = torch.randint(0, 1 << 22, (n,), dtype=torch.int32, device="cuda", generator=g)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    wk, order = torch.sort(keys, stable=True)
    k, v = keys.clone(), vals.clone()
    gpu_engine.sort_pairs_dev(k, v, n, 22)
    torch.cuda.synchronize()
    assert torch.equal(k, wk) and torch.equal(v, vals[order])
