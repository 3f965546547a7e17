"""BASELINE.json config 5: one 2^LOG-point G1 MSM point-range-sharded over the ranks of a torchrun launch (strong
scaling: the total is fixed, each of N ranks owns 2^LOG / N points with its own SRS shard + window table), NCCL
all-gather of the 192-byte partials, host fold.  Also runs at N = 1 (plain `python scripts/msm_sharded.py 26`).
Seeds do not depend on the rank: every N computes the same point (`result_sha256`; bench.py reports the same workload
in `extra.multi_gpu.g1_msm_2p26_strong`).  Prints one JSON line: device time per MSM (CUDA events, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402


def main():
    log_total = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = z.Engine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    # the SAME global problem whatever the number of ranks: rank g takes points / scalars [lo, hi) of one seeded set
    total = 1 << log_total
    lo, hi = z.dist.shard_range(total, rank, world)
    n = hi - lo
    bases = torch.zeros(n * 12, dtype=torch.int64, device=dev)
    eng.generate_bases_dev(0x2627, n, bases, first=lo)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x2626)
    allsc = torch.randint(0, 2**62, (total * 4,), dtype=torch.int64, device=dev, generator=gen)
    scalars = allsc if world == 1 else allsc[lo * 4:hi * 4].clone()
    del allsc
    eng.srs_upload_dev(bases, n)
    del bases
    torch.cuda.empty_cache()
    eng.srs_precompute()
    torch.cuda.synchronize()

    def step():
        if world == 1:
            return eng.msm_dev(scalars, None, n)
        return z.dist.msm_sharded(eng, scalars, None, n, device=dev)

    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        c, w = eng.last_msm_shape()
        print(json.dumps({"op": "msm_sharded", "log_total": log_total, "n_gpus": world, "points_per_gpu": n, "ms": ms,
                          "c": c, "windows": w, "mpoints_per_s": (1 << log_total) / ms / 1e3,
                          "result_sha256": __import__("hashlib").sha256(out[0].tobytes()).hexdigest()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
