#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "G1 MSM ms at 2^20/2^24, Fr NTT ms
at 2^24 ...; 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

One "step" = one 2^24-point G1 MSM on one batch of synthetic input: random scalars against a resident
SRS of 2^24 distinct pseudo-random points generated on the device, with the fixed-base window table
built once at setup (zkp_srs_precompute, the KzgScheme::new side of the seam; not in the timed region,
exactly as the SRS upload is not).  At N > 1 every rank owns a 2^24-point shard of an N * 2^24-point
MSM (point-range sharding, weak scaling): local Pippenger, NCCL all-gather of the 192-byte partials,
host fold.  The JSON line's `value` is device time per step with inputs resident
in HBM; `e2e` is the same step through the reference-facing entry point `zkp_msm_g1` with the
scalars in pinned HOST memory (H2D copy and result read-back inside the timed region).  `extra`
carries the other numbers BASELINE.json names (MSM 2^20, NTT 2^24) measured the same way.

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (as the thing
timed on the CPU, never on the product path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N_MSM = 24
LOG_N_NTT = 24
IMAD_PER_FQ_MUL = 300       # 12-limb CIOS = 2*12^2 + 12 multiply-adds
FQ_MUL_PER_MADD = 10        # XYZZ mixed add 8M + 2S
IMAD_PER_FR_MUL = 136       # 8-limb CIOS = 2*8^2 + 8
WORKLOAD = ("G1 MSM, 2^24 points per GPU (BLS12-381), random scalars < 2^254, resident SRS of distinct generated "
            "points + fixed-base window table; N GPUs = point-range shards of an N*2^24-point MSM + NCCL "
            "all-gather of partials")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, name in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        busy = [s for s in sm if s > 500] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm on the box's host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(log_sample: int = 12):
    """kzg/src/scheme.rs:84-96 restated literally (oracle/zkp_oracle.c:orc_msm_naive): per-term
    double-and-add + into_affine + affine fold, single-threaded like the reference (no rayon, no
    `parallel` feature).  Timed on 2^log_sample terms and scaled linearly to 2^24 (the algorithm is
    a serial loop of independent, equal-cost terms)."""
    import numpy as np
    import zkp_implementation_b200 as z
    from oracle import coracle as c

    c.build()
    F = z.fields
    n = 1 << log_sample
    bases = c.srs(F.fr_to_mont_array([0xB200]), 64)
    bases = np.tile(bases, (n // 64, 1))
    s = F.random_fr_mont(0xC0DE, n)
    t0 = time.perf_counter()
    c.msm_naive(s, bases)
    dt = time.perf_counter() - t0
    ms_full = dt * 1e3 * (1 << LOG_N_MSM) / n
    return ms_full, dt, n


def cpu_best_effort(log_sample: int = 20):
    """Pippenger on all host cores (oracle/zkp_oracle.c:orc_msm_pippenger) -- what a tuned CPU port
    would do; reported beside the faithful line, never instead of it."""
    import numpy as np
    import zkp_implementation_b200 as z
    from oracle import coracle as c

    F = z.fields
    n = 1 << log_sample
    bases = np.tile(c.srs(F.fr_to_mont_array([0xB200]), 256), (n // 256, 1))
    s = F.random_fr_mont(0xC0DF, n)
    t0 = time.perf_counter()
    c.msm_pippenger(s, bases)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    c.ntt(F.random_fr_mont(3, 1 << 22), 22)
    dn = time.perf_counter() - t1
    return {"msm_pippenger_ms_scaled_2p24": dt * 1e3 * (1 << LOG_N_MSM) / n, "msm_sample": "2^%d points, %.2f s" % (log_sample, dt),
            "ntt_radix2_ms_scaled_2p24": dn * 1e3 * 4 * 24 / 22, "ntt_sample": "2^22 points, %.2f s" % dn,
            "cores": c.num_threads()}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        ms_full, dt, n = cpu_reference_sample(11)
        if i >= args.warmup:
            vals.append(ms_full)
    v = statistics.median(vals)
    sample = "evaluate_in_s restated (per-term scalar mul + into_affine + affine fold) on 2^11 terms per step, scaled x2^13 to 2^24"
    line = {
        "impl": "reference", "metric": "g1_msm_2p24_ms", "value": v, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (Fq 6x64, Fr 4x64 Montgomery)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "reference algorithm kzg/src/scheme.rs:84-96 restated, host CPU"},
        "cpu_baseline": {"value": v, "unit": "ms", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------------
def run_engine(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    import zkp_implementation_b200 as z

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = z.Engine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_profiling(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, on_start=None):
        """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        if on_start:
            on_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    n = 1 << LOG_N_MSM
    # ---- synthetic inputs, resident in HBM (each rank its own shard: seed depends on rank) ----
    bases = torch.zeros(n * 12, dtype=torch.int64, device=dev)
    eng.generate_bases_dev(0xB200 + rank, n, bases)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED + rank)
    scalars = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device=dev, generator=gen)  # < 2^254 < r
    eng.srs_upload_dev(bases, n)  # resident SRS for the reference-facing entry point
    t_tab = time.perf_counter()
    eng.srs_precompute()          # fixed-base window table (setup, next to KzgScheme::new)
    torch.cuda.synchronize()
    t_tab = time.perf_counter() - t_tab
    host_scalars = torch.empty(n * 4, dtype=torch.int64, pin_memory=True)
    host_scalars.copy_(scalars)
    host_np = host_scalars.numpy().view(np.uint64)
    torch.cuda.synchronize()

    imad_wide, imad_lo = eng.imad_peak()
    launches = {"n": 0}
    phase_acc = {"accumulate": [], "sort": [], "recode": [], "bounds_tasks": [], "reduce": []}
    result = {}

    def step_resident():
        if world == 1:
            out, inf = eng.msm_dev(scalars, None, n)  # bases = resident SRS
        else:
            out, inf = z.dist.msm_sharded(eng, scalars, None, n, device=dev)
        result["resident"] = out
        launches["n"] += eng.last_launches("msm")
        for k, v in eng.last_phase_ms().items():
            phase_acc[k].append(v)

    def step_e2e():
        if world == 1:
            out, inf = eng.msm(host_np)  # zkp_msm_g1: H2D of 2^24 x 32 B, Pippenger, affine result to host
        else:
            sdev = torch.empty_like(scalars)
            sdev.copy_(host_scalars, non_blocking=True)
            out, inf = z.dist.msm_sharded(eng, sdev, None, n, device=dev)
        result["e2e"] = out

    # warm-up also sizes every scratch buffer
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for k in phase_acc:
        phase_acc[k].clear()
    def reset_counters():
        launches["n"] = 0
        for k in phase_acc:
            phase_acc[k].clear()

    ms = timed(step_resident, args.steps, args.warmup, on_start=reset_counters)
    timed_launches = launches["n"]
    acc_ms = statistics.mean(phase_acc["accumulate"][-args.steps:])
    phases = {k: statistics.mean(v[-args.steps:]) for k, v in phase_acc.items()}
    clocks = sampler.stop() if sampler else None
    e2e_ms = timed(step_e2e, args.steps, max(1, args.warmup - 1))
    assert (result["resident"] == result["e2e"]).all(), "resident and host-buffer paths disagree"
    c_bits, n_win = eng.last_msm_shape()

    extra = {}
    if rank == 0:
        # ---- the other numbers BASELINE.json names, N = 1 semantics on rank 0's GPU ----
        n20 = 1 << 20
        extra["g1_msm_2p24_srs_table_build_s"] = t_tab
        extra["g1_msm_2p24_adhoc_bases_ms"] = timed_local(torch, lambda: eng.msm_dev(scalars, bases, n), 3, 1)
        extra["g1_msm_2p20_adhoc_bases_ms"] = timed_local(torch, lambda: eng.msm_dev(scalars, bases, n20), 5, 3)
        nt = 1 << LOG_N_NTT
        poly = torch.randint(0, 2**62, (nt * 4,), dtype=torch.int64, device=dev)
        ntt_ms = timed_local(torch, lambda: eng.ntt_dev(poly, LOG_N_NTT), 10, 3)
        extra["fr_ntt_2p24_ms"] = ntt_ms
        extra["fr_intt_2p24_ms"] = timed_local(torch, lambda: eng.ntt_dev(poly, LOG_N_NTT, inverse=True), 5, 2)
        extra["fr_ntt_2p24_launches"] = eng.last_launches("ntt")
        hp = torch.empty(nt * 4, dtype=torch.int64, pin_memory=True)
        hp.copy_(poly)
        hp_np = hp.numpy().view(np.uint64)
        extra["fr_ntt_2p24_e2e_ms"] = timed_local(torch, lambda: eng.ntt(hp_np, LOG_N_NTT), 3, 1)
        hbm, hbm_src = peaks()
        passes_min = -(-LOG_N_NTT // 12)
        ntt_bytes = 64.0 * nt * passes_min
        ntt_imad = (nt // 2) * LOG_N_NTT * IMAD_PER_FR_MUL
        extra["ntt_roofline"] = {
            "bound": "hbm", "achieved": ntt_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": ntt_bytes / (ntt_ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_src,
            "algorithmic_bytes": ntt_bytes, "passes_counted": passes_min, "passes_run": eng.last_launches("ntt"),
            "traffic": ntt_traffic(eng.last_launches("ntt")),
            "integer_bound_note": "butterfly arithmetic = N/2*log2(N)*136 IMAD; fraction of the measured IMAD.WIDE peak below",
            "int_pipe_frac": ntt_imad / (ntt_ms * 1e-3) / imad_wide if imad_wide else None,
        }
        del poly, hp
        # ---- SRS of 2^20 + 3 points: the 2^20 commitment and the end-to-end PLONK prove (config 4) ----
        if world == 1:
            extra.update(plonk_and_2p20(z, eng, torch, scalars))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (msm_accumulate_kernel), integer pipe ----
    alg_imad = float(IMAD_PER_FQ_MUL * FQ_MUL_PER_MADD) * n * n_win  # 3000 IMAD per (point, window) mixed add
    achieved = alg_imad / (acc_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "msm_accumulate_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {
        "bound": "int-pipe", "kernel": "msm_accumulate_kernel", "achieved": achieved, "peak": imad_wide / 1e12,
        "unit": "TIMAD/s", "frac": achieved / (imad_wide / 1e12) if imad_wide else None,
        "peak_source": "measured in this run: IMAD.WIDE.U32.X carry chains on all SMs (zkp_bench_imad_peak)",
        "peak_imad_lo": imad_lo / 1e12, "frac_of_imad_lo_peak": achieved / (imad_lo / 1e12) if imad_lo else None,
        "algorithmic_imad_per_launch": alg_imad, "kernel_ms": acc_ms, "kernel_share_of_step": acc_ms / ms,
        "window_bits": c_bits, "windows": n_win, "traffic": traffic, "phases_ms": phases,
    }
    cpu = None
    if world == 1:  # the CPU baseline is an N = 1 measurement (torchrun pins OMP_NUM_THREADS = 1)
        ref_ms, ref_dt, ref_n = cpu_reference_sample(12)
        cpu = {"value": ref_ms, "unit": "ms", "cores": 1, "kind": "port",
               "sample": "evaluate_in_s restated (oracle/zkp_oracle.c:orc_msm_naive) on 2^12 terms = %.2f s, scaled x2^12 to "
                         "2^24; single-threaded like the reference" % ref_dt,
               "best_effort_all_cores": cpu_best_effort(20)}
    line = {
        "metric": "g1_msm_2p24_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (Fq 12x32, Fr 8x32 Montgomery)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "points_total": n * world, "window_bits": c_bits, "windows": n_win,
                   "l2": "inputs (2 GiB bases+scalars, 2 GiB digit pairs) exceed the 126 MB L2; no explicit flush"},
        "points_per_s": n * world / (ms * 1e-3),
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 97,
                "api": "zkp_msm_g1 (pinned host scalars -> affine point on host; SRS resident)"},
        "gpu_launches": timed_launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def plonk_and_2p20(z, eng, torch, scalars):
    """BASELINE.json config 4 on this GPU: synthetic chain circuit of 2^20 - 3 gates, SRS from a fixed secret with
    its window table, fixed blinding; wall clock inside zkp_plonk_prove (all rounds, transcript included)."""
    import hashlib

    k = 20
    n = 1 << k
    secret = 0x1F2E3D4C5B6A79881234567
    blind = [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]
    eng.srs_generate(secret, n + 3, want_points=False)
    eng.srs_precompute()
    out = {"g1_msm_2p20_ms": timed_local(torch, lambda: eng.msm_dev(scalars, None, n), 5, 3)}
    cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
    runs = []
    digest = None
    for _ in range(4):
        p = z.plonk.generate_proof(cc, blind)
        runs.append(p.timings_ms["total"])
        digest = hashlib.sha256(p.to_bytes()).hexdigest()
    cc.close()
    out["plonk_prove_2p20_ms"] = min(runs[1:])
    out["plonk_prove_2p20_runs_ms"] = runs
    out["plonk_proof_sha256"] = digest
    out["plonk_config"] = "chain circuit, 2^20 - 3 gates (alternating mul / add, c_i wired to a_(i+1)), seed 20, fixed b1..b9"
    return out


def ntt_traffic(passes):
    """DRAM bytes of one 2^24 transform = passes x the per-launch figure of the committed ncu capture."""
    tp = os.path.join(ROOT, "profiles", "ntt_pass_traffic.json")
    if not os.path.exists(tp):
        return None
    return json.load(open(tp)).get("dram_bytes_per_launch", 0) * passes


def timed_local(torch, fn, steps, warmup):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
