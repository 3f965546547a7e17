"""BASELINE.json config 3, batched case over several GPUs: a batch of independent prover polynomials shards whole
across the ranks (dist.batch_shard: polynomial i -> rank i % world), no collective on the data path.  Prints the
time of the slowest rank for the whole batch (CUDA events, max over ranks via one all-reduce of the timing)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = z.Engine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for log_n, batch in ((20, 32), (22, 32), (24, 16), (26, 8)):
        mine = list(z.dist.batch_shard(batch, rank, world))
        n = 1 << log_n
        data = torch.randint(0, 2**62, (max(len(mine), 1) * n * 4,), dtype=torch.int64, device=dev)
        run = lambda: eng.ntt_dev(data, log_n, batch=len(mine)) if mine else None
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank == 0:
            print(json.dumps({"op": "ntt_batch_sharded", "log_n": log_n, "batch": batch, "n_gpus": world, "ms": ms,
                              "polys_per_gpu": len(mine), "gelem_per_s": batch * n / ms / 1e6}), flush=True)
        del data
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
