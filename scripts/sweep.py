"""BASELINE.json configs 2 and 3 on one B200: G1 MSM sweep 2^10..2^24 (random scalars, distinct generated
bases resident in HBM, per-phase device times) and Fr NTT / iNTT / coset sweep 2^16..2^26, single and batched.
One JSON line per measurement; CUDA events on the engine's stream, 3 warm-up + K timed runs."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = torch.device("cuda", 0)
    eng = z.Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_profiling(True)
    if what in ("all", "msm"):
        nmax = 1 << 24
        bases = torch.zeros(nmax * 12, dtype=torch.int64, device=dev)
        eng.generate_bases_dev(0xB200, nmax, bases)
        scalars = torch.randint(0, 2**62, (nmax * 4,), dtype=torch.int64, device=dev)
        eng.srs_upload_dev(bases, nmax)
        for log_n in (10, 12, 14, 16, 18, 20, 22, 24):
            n = 1 << log_n
            steps = 10 if log_n <= 20 else 4
            ms = timed(lambda: eng.msm_dev(scalars, bases, n), steps)
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.msm_dev(scalars, bases, n)
            wall = (time.perf_counter() - t0) * 1e3 / steps
            c, w = eng.last_msm_shape()
            print(json.dumps({"op": "msm", "log_n": log_n, "ms": ms, "wall_ms": wall, "c": c, "windows": w,
                              "phases_ms": eng.last_phase_ms(), "launches": eng.last_launches("msm"),
                              "mpoints_per_s": n / ms / 1e3}), flush=True)
        # fixed-base path: the same points as a resident SRS with the precomputed window table
        for log_n in (16, 18, 20, 22, 24):
            n = 1 << log_n
            eng.srs_upload_dev(bases, n)
            for bits in ((0,) if log_n < 24 else (0, 22, 23)):
                t0 = time.perf_counter()
                eng.srs_precompute(bits)
                torch.cuda.synchronize()
                t_tab = time.perf_counter() - t0
                steps = 10 if log_n <= 20 else 4
                ms = timed(lambda: eng.msm_dev(scalars, None, n), steps)
                c, w = eng.last_msm_shape()
                ref = eng.msm_dev(scalars, bases, n)[0]
                same = bool((eng.msm_dev(scalars, None, n)[0] == ref).all())
                print(json.dumps({"op": "msm_fixed_base", "log_n": log_n, "ms": ms, "c": c, "windows": w,
                                  "table_build_s": t_tab, "table_gib": w * n * 96 / 2**30, "matches_windowed": same,
                                  "phases_ms": eng.last_phase_ms(), "launches": eng.last_launches("msm"),
                                  "mpoints_per_s": n / ms / 1e3}), flush=True)
        eng.srs_upload_dev(bases, 1)
        del bases, scalars
        torch.cuda.empty_cache()
    if what in ("all", "ntt"):
        for log_n in (16, 18, 20, 22, 24, 26):
            n = 1 << log_n
            for batch in (1, 8, 32):
                if (n * batch * 32) > 40 * 2**30:
                    continue
                data = torch.randint(0, 2**62, (n * batch * 4,), dtype=torch.int64, device=dev)
                steps = 10 if n * batch <= 1 << 24 else 3
                row = {"op": "ntt", "log_n": log_n, "batch": batch}
                row["fwd_ms"] = timed(lambda: eng.ntt_dev(data, log_n, batch), steps)
                row["launches"] = eng.last_launches("ntt")
                if batch == 1:
                    row["inv_ms"] = timed(lambda: eng.ntt_dev(data, log_n, batch, inverse=True), steps)
                    row["coset_fwd_ms"] = timed(lambda: eng.ntt_dev(data, log_n, batch, coset=7), steps)
                row["melem_per_s"] = n * batch / row["fwd_ms"] / 1e3
                row["hbm_frac_2pass"] = 64.0 * n * batch * (-(-log_n // 12)) / (row["fwd_ms"] * 1e-3) / 1e9 / 6546.2
                print(json.dumps(row), flush=True)
                del data
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
