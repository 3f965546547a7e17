"""Single-process multi-GPU context (zkp_ctx_create_multi): one `KzgScheme`-shaped handle over several devices, the
resident SRS sharded by point range, every commitment folded from per-device partial sums -- no torch.distributed in
between (kzg/src/scheme.rs:34,49-52: one value, `&self`).  On a box with fewer GPUs than shards the device id repeats
(several shards on one GPU / on the emulator): the sharding logic is identical."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _multi_engine(zkp, kind, shards):
    if kind == "emu":
        import importlib.util

        spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return zkp.Engine(lib_path=mod.build_emu(), devices=[0] * shards)
    import torch

    assert torch.cuda.is_available()
    ndev = torch.cuda.device_count()
    return zkp.Engine(devices=[g % ndev for g in range(shards)])


@pytest.mark.parametrize("kind", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("shards", [2, 3])
def test_multi_ctx_commitments(zkp, coracle, kind, shards):
    F = zkp.fields
    eng = _multi_engine(zkp, kind, shards)
    try:
        assert eng.shards() == shards
        n = 1000 if kind == "emu" else 50000
        pts = eng.srs_generate(0xFEED, n)  # every shard generates its own point range
        assert eng.srs_len() == n
        small = 40
        assert (pts[:small] == coracle.srs(F.fr_to_mont_array([0xFEED]), small)).all()
        s = F.random_fr_mont(0x3C, n)
        want = coracle.msm_pippenger(s, pts)
        for table in (False, True):
            if table:
                eng.srs_precompute()
            out, inf = eng.msm(s)  # zkp_msm_g1: host scalars, every shard takes its slice
            assert (out == want).all() and not inf, table
            # prefixes: shorter than one shard, ending inside a shard, the empty sum
            for m in (1, n // shards - 1, n // 2 + 7):
                assert (eng.msm(s[:m])[0] == coracle.msm_pippenger(s[:m], pts[:m])).all(), (table, m)
            assert eng.msm(s[:0])[1]
            # device-resident scalars on the primary device (zkp_msm_g1_dev / zkp_msm_g1_multi_dev)
            dv = [eng.vec(s), eng.vec(s[:n - 5])]
            out2, _ = eng.msm_dev(dv[0].ptr, None, n)
            assert (out2 == want).all()
            got = eng.msm_multi_dev([d.ptr for d in dv], [n, n - 5])
            assert (got[0][0] == want).all() and (got[1][0] == coracle.msm_pippenger(s[:n - 5], pts[:n - 5])).all()
        # uploaded SRS (zkp_srs_upload) is split the same way; too long a polynomial is still the reference's panic
        eng.srs_upload(pts[:777])
        assert (eng.msm(s[:777])[0] == coracle.msm_pippenger(s[:777], pts[:777])).all()
        with pytest.raises(zkp.ZkpError) as ei:
            eng.msm(s[:778])
        assert ei.value.status == 4
    finally:
        eng.close()


@pytest.mark.parametrize("kind", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_multi_ctx_plonk_golden_proof(zkp, kind):
    """The whole prover over a multi context (transforms on the primary device, commitments sharded): golden bytes."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk.json")))
    secret, blind = int(gold["secret"], 16), [int(b, 16) for b in gold["blinding"]]
    eng = _multi_engine(zkp, kind, 2)
    try:
        circ = zkp.plonk.Circuit()
        circ.add_multiplication_gate((0, 0, 1), (1, 0, 2), (0, 1, 2), 0)
        circ.add_multiplication_gate((2, 0, 2), (1, 1, 3), (2, 1, 6), 0)
        zkp.KzgScheme(eng, zkp.Srs.new_from_secret(eng, secret, 2))
        cc = circ.compile(eng)
        want = gold["circuits"]["circuit_accepted_03"]["proof"]
        assert zkp.plonk.generate_proof(cc, blind).to_bytes().hex() == want
        assert zkp.plonk.generate_proof(cc, blind, products=True).to_bytes().hex() == want
        cc.close()
    finally:
        eng.close()
