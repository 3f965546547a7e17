"""Multi-process host logic (world_size 2, gloo, CPU): point-range sharding + all-gather of partials
+ host fold, on the kernel emulator.  The same code path runs over NCCL on GPUs (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_path, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import zkp_implementation_b200 as z
    from oracle import coracle as c

    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = z.fields
    eng = z.Engine(0, lib_path=emu_path)
    s = F.random_fr_mont(123, n)
    b = c.srs(F.fr_to_mont_array([77]), n)
    lo, hi = z.dist.shard_range(n, rank, world)
    sl, bl = np.ascontiguousarray(s[lo:hi]), np.ascontiguousarray(b[lo:hi])
    out, inf = z.dist.msm_sharded(eng, sl, bl, hi - lo)  # emulator: "device" pointers are host arrays
    exp = c.msm_pippenger(s, b)
    # whole-polynomial batch sharding of NTTs: rank r transforms polynomials r, r+world, ...
    batch, log_n = 3, 8
    polys = F.random_fr_mont(5, batch << log_n).reshape(batch, -1, 4)
    ok_ntt = True
    for i in z.dist.batch_shard(batch, rank, world):
        d = polys[i].copy()
        eng.ntt(d, log_n)
        ok_ntt &= bool((d == c.ntt(polys[i], log_n)).all())
    q.put((rank, bool((out == exp).all()) and not inf, ok_ntt))
    dist.destroy_process_group()


def test_sharded_msm_gloo_world2(zkp):
    import importlib.util
    import torch.multiprocessing as mp

    from oracle import coracle
    coracle.build()
    spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    emu = mod.build_emu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu, 301, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def test_shard_ranges(zkp):
    for n in (0, 1, 7, 8, 1 << 20):
        for w in (1, 2, 3, 8):
            rs = [zkp.dist.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def _ntt_worker(rank, world, port, emu_path, log_n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import zkp_implementation_b200 as z
    from oracle import coracle as c

    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = z.fields
    eng = z.Engine(0, lib_path=emu_path)
    n = 1 << log_n
    x = F.random_fr_mont(900 + log_n, n)
    d = z.dist.DistNtt(eng, log_n, rank, world)
    ok = []
    for coset in (None, 7):
        cs = None if coset is None else F.fr_to_mont_array([coset])[0]
        want = c.ntt(x, log_n, coset_mont=cs) if coset else c.ntt(x, log_n)
        a = torch.from_numpy(d.layout_a(x).view(np.int64).reshape(-1).copy())
        b = d.forward(a, coset=coset)
        ok.append(bool((b.numpy().view(np.uint64).reshape(-1, 4) == d.layout_b(want)).all()))
        back = d.inverse(b, coset=coset)
        ok.append(bool((back.numpy().view(np.uint64).reshape(-1, 4) == d.layout_a(x)).all()))
    q.put((rank, all(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 6), (2, 13), (4, 9)])
def test_four_step_ntt_gloo(zkp, world, log_n):
    """Distributed four-step NTT (stage kernel + all-to-all + local batched NTT) == single NTT of the oracle,
    forward / inverse / coset, on the kernel emulator with gloo."""
    import importlib.util
    import torch.multiprocessing as mp

    from oracle import coracle
    coracle.build()
    spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    emu = mod.build_emu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + world * 7 + log_n) % 2000
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, emu, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _plonk_worker(rank, world, port, emu_path, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import json
    import torch.distributed as dist
    import zkp_implementation_b200 as z
    from oracle import plonk_ref as ref

    dist.init_process_group("gloo", rank=rank, world_size=world)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk.json")))
    secret, blind = int(gold["secret"], 16), [int(b, 16) for b in gold["blinding"]]
    eng = z.Engine(0, lib_path=emu_path)
    ok = True
    for name in ("circuit_accepted_02", "circuit_accepted_03"):
        rc = getattr(ref, name)()
        c = z.plonk.Circuit()
        for i, g in enumerate(rc.gates):
            wires = [(pos[0], pos[1], rc.vals[col][i]) for col, pos in enumerate((g.a, g.b, g.c))]
            add = c.add_multiplication_gate if g.q_m == 1 else (c.add_addition_gate if g.q_r == 1 else c.add_constant_gate)
            add(*wires, (-g.pi) % ref.R)
        n = gold["circuits"][name]["size"]
        total = n + 3
        lo, hi = z.dist.shard_range(total, rank, world)
        eng.srs_generate(secret, hi - lo, want_points=False, first=lo)  # this rank's point range of the SRS
        eng.srs_precompute()
        cc = c.compile(eng)
        proof = z.plonk.generate_proof_sharded(cc, blind, rank, world, lo, total)
        ok &= proof.to_bytes().hex() == gold["circuits"][name]["proof"]
        cc.close()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prover_gloo(zkp, world):
    """zkp_plonk_prove_sharded: every rank holds a point range of the SRS, commits its range of each polynomial, the
    192-byte partials are all-gathered and folded -- every rank returns the golden proof bytes."""
    import importlib.util
    import torch.multiprocessing as mp

    spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    emu = mod.build_emu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + world * 11) % 2000
    procs = [ctx.Process(target=_plonk_worker, args=(r, world, port, emu, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
