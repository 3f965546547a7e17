"""Multi-GPU four-step NTT on real GPUs (run under torchrun, one rank per GPU): parity of both transports
(NCCL all-to-all, fused peer-memory exchange) against the single-GPU transform of the same vector, then
device timings (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        scripts/dist_ntt_check.py 20 24 26
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = z.Engine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    logs = [int(a) for a in sys.argv[1:]] or [16, 22]
    for log_n in logs:
        n = 1 << log_n
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + log_n)  # same vector on every rank
        x = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device=dev, generator=gen)
        want = x.clone()
        eng.ntt_dev(want, log_n)  # single-GPU reference of the same engine (itself oracle-checked)
        res = {"log_n": log_n, "world": world}
        for p2p in (False, True):
            try:
                d = z.dist.DistNtt(eng, log_n, rank, world, p2p=p2p)
            except Exception as ex:  # peer access unavailable on this box
                res["p2p_error" if p2p else "nccl_error"] = repr(ex)[:200]
                continue
            cols = (1 << d.s) // world
            a = x.view(1 << d.r, 1 << d.s, 4)[:, rank * cols:(rank + 1) * cols, :].contiguous().view(-1)
            wb = want.view(1 << d.s, 1 << d.r, 4)[:, rank * d.rows_local:(rank + 1) * d.rows_local, :]
            wb = wb.permute(1, 0, 2).contiguous().view(-1)
            a0 = a.clone()
            b = d.forward(a.clone())
            ok_f = bool(torch.equal(b, wb))
            back = d.inverse(b.clone())
            ok_i = bool(torch.equal(back, a0))
            oks = torch.tensor([int(ok_f), int(ok_i)], device=dev)
            dist.all_reduce(oks, op=dist.ReduceOp.MIN)
            tag = "p2p" if p2p else "nccl"
            res[tag + "_forward_ok"], res[tag + "_inverse_ok"] = bool(oks[0].item()), bool(oks[1].item())

            def timed(fn, steps=5, warmup=2):
                for _ in range(warmup):
                    fn()
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record()
                torch.cuda.synchronize(); dist.barrier()
                t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())

            work = a0.clone()
            if p2p:
                res[tag + "_forward_ms"] = timed(lambda: d.forward_into_exchange(work))
                out = torch.empty_like(work)
                res[tag + "_inverse_ms"] = timed(lambda: d.inverse_from_exchange(out))
            else:
                res[tag + "_forward_ms"] = timed(lambda: d.forward(work))
                res[tag + "_inverse_ms"] = timed(lambda: d.inverse(work))
            d.close()
        single = x.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.ntt_dev(single, log_n)
        e0.record()
        for _ in range(3):
            eng.ntt_dev(single, log_n)
        e1.record()
        torch.cuda.synchronize()
        res["single_gpu_ms"] = e0.elapsed_time(e1) / 3
        if rank == 0:
            print(json.dumps(res), flush=True)
        del x, want, single
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
