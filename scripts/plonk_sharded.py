"""BASELINE.json config 4 over several GPUs (torchrun, one rank per GPU): replicated prover, point-range-sharded
commitments (zkp_plonk_prove_sharded).  Prints one JSON line per size with the time of the slowest rank and the
SHA-256 of the proof (every rank must produce the same bytes as the single-GPU prover)."""
import hashlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

SECRET = 0x1F2E3D4C5B6A79881234567
BLIND = [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = z.Engine(local)
    for k in [int(a) for a in sys.argv[1:]] or [16, 20]:
        n = 1 << k
        total = n + 3
        lo, hi = z.dist.shard_range(total, rank, world)
        eng.srs_generate(SECRET, hi - lo, want_points=False, first=lo)
        eng.srs_precompute()
        cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
        runs, digest = [], None
        for _ in range(4):
            if world > 1:
                torch.cuda.synchronize()
                dist.barrier()
            t0 = time.perf_counter()
            if world > 1:
                p = z.plonk.generate_proof_sharded(cc, BLIND, rank, world, lo, total, device=dev)
            else:
                p = z.plonk.generate_proof(cc, BLIND)
            runs.append((time.perf_counter() - t0) * 1e3)
            digest = hashlib.sha256(p.to_bytes()).hexdigest()
        best = min(runs[1:])
        if world > 1:
            t = torch.tensor([best], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item())
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            assert len(set(digests)) == 1, digests
        if rank == 0:
            print(json.dumps({"op": "plonk_prove_sharded", "log_n": k, "n_gpus": world, "prove_ms": best,
                              "inside_prove_ms": p.timings_ms, "proof_sha256": digest}), flush=True)
        cc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
