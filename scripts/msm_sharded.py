"""BASELINE.json config 5: one 2^LOG-point G1 MSM point-range-sharded over the ranks of a torchrun launch (strong
scaling: the total is fixed, each of N ranks owns 2^LOG / N points with its own SRS shard + window table), NCCL
all-gather of the 192-byte partials, host fold.  Also runs at N = 1 (plain `python scripts/msm_sharded.py 26`).
Prints one JSON line: device time per MSM (CUDA events, max over ranks)."""
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
    n = (1 << log_total) // world
    bases = torch.zeros(n * 12, dtype=torch.int64, device=dev)
    eng.generate_bases_dev(0xB200 + rank, n, bases)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED + rank)
    scalars = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device=dev, generator=gen)
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
                          "result_x0": hex(int(out[0][0]))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
