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
carries the other numbers BASELINE.json names (MSM 2^20, NTT 2^24, PLONK prove) measured the same way, the
host-CPU baselines of the reference's own algorithms (per-term MSM, single-core NTT, the O(n^2)-accumulator
prover) and -- at N > 1, on all ranks -- the other multi-GPU rows of SURVEY.md 8(e): the strong-scaling
2^26 MSM (config 5), the four-step 2^26 NTT with both exchange transports, whole-polynomial batches and the
sharded 2^20 prover.

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
FQ_MUL_PER_AFFINE_ADD_KERNEL = 5 - 2 / 16  # batched-affine addition kernel: lambda, lambda^2, lambda*dx + 2 back-substitution
                                          # products, the latter skipped for one of the K = 16 outputs of a thread
IMAD_PER_FR_MUL = 136       # 8-limb CIOS = 2*8^2 + 8
IMAD_PER_FR_MUL_SASS = 121  # what fp_mul<Fr> executes: 110 IMAD.WIDE + 11 IMAD (r = 1 mod 2^32 shortens the reduction rows)
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


def host_threads() -> int:
    """Host cores this process may use (torchrun pins OMP_NUM_THREADS = 1 per rank: the CPU legs run on rank 0 only and
    ask for the cores explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_best_effort(log_sample: int = 20):
    """Pippenger on all host cores (oracle/zkp_oracle.c:orc_msm_pippenger) -- what a tuned CPU port
    would do; reported beside the faithful line, never instead of it.  Plus the NTT both ways: all cores, and ONE core
    (the reference is single-threaded) at the metric's own size."""
    import numpy as np
    import zkp_implementation_b200 as z
    from oracle import coracle as c

    F = z.fields
    th = host_threads()
    n = 1 << log_sample
    bases = np.tile(c.srs(F.fr_to_mont_array([0xB200]), 256), (n // 256, 1))
    s = F.random_fr_mont(0xC0DF, n)
    t0 = time.perf_counter()
    c.msm_pippenger(s, bases, threads=th)
    dt = time.perf_counter() - t0
    a22 = F.random_fr_mont(3, 1 << 22)
    t1 = time.perf_counter()
    c.ntt(a22, 22, threads=th)
    dn = time.perf_counter() - t1
    a24 = F.random_fr_mont(4, 1 << LOG_N_NTT)
    t2 = time.perf_counter()
    c.ntt(a24, LOG_N_NTT, threads=1)
    d1 = time.perf_counter() - t2
    return {"msm_pippenger_ms_scaled_2p24": dt * 1e3 * (1 << LOG_N_MSM) / n, "msm_sample": "2^%d points, %.2f s" % (log_sample, dt),
            "ntt_radix2_ms_scaled_2p24": dn * 1e3 * 4 * 24 / 22, "ntt_sample": "2^22 points, %.2f s" % dn,
            "ntt_radix2_single_core_2p24_ms": d1 * 1e3, "ntt_single_core_sample": "2^24 points (full size), 1 thread",
            "cores": th}


PLONK_SECRET = 0x1F2E3D4C5B6A79881234567


def plonk_blinding(z):
    return [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]


def cpu_plonk_baseline(z, gpu_eng):
    """The reference's prover on host cores: host/plonk.cpp linked against the CPU backend of the
    C ABI (oracle/cpu_backend.cpp) -- the same orchestration, the oracle's kernels.
      * `reference`: single-threaded, per-term MSM (scheme.rs:84-96), one product per `&a * &b`, and the O(n^2)
        `compute_acc` of prover.rs:302-377, at n = 2^10 and (when 2^10 took < 6 s) 2^12;
      * `all_cores`: the same prover with the O(n) accumulator, Pippenger MSM and threaded NTTs, at n = 2^16 and 2^20.
    The SRS is generated on the GPU and handed over (setup is not what is being compared)."""
    import hashlib

    import numpy as np
    from oracle.cpu_engine import CpuEngine

    blind = plonk_blinding(z)
    out = {"reference_single_core": {}, "all_cores": {}, "cores": host_threads()}

    def prove(eng, k, **kw):
        n = 1 << k
        eng.srs_upload(gpu_eng.srs_generate(PLONK_SECRET, n + 3, want_points=True))
        cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
        t0 = time.perf_counter()
        p = z.plonk.generate_proof(cc, blind, **kw)
        ms = (time.perf_counter() - t0) * 1e3
        cc.close()
        return {"prove_ms": ms, "inside_ms": p.timings_ms, "proof_sha256": hashlib.sha256(p.to_bytes()).hexdigest()}

    eng = CpuEngine(threads=1, pippenger=False)
    r10 = prove(eng, 10, reference_acc=True)
    out["reference_single_core"]["2^10"] = r10
    if r10["prove_ms"] < 6000:
        out["reference_single_core"]["2^12"] = prove(eng, 12, reference_acc=True)
    eng.close()
    eng = CpuEngine(threads=host_threads(), pippenger=True)
    for k in (16, 20):
        out["all_cores"]["2^%d" % k] = prove(eng, k, products=True)
        if out["all_cores"]["2^%d" % k]["prove_ms"] > 60000:
            break
    eng.close()
    out["note"] = ("reference_single_core = prover.rs as written (O(n^2) compute_acc, per-term MSM, 1 thread); all_cores = same "
                   "orchestration with the O(n) accumulator, Pippenger MSM and threaded NTTs")
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        ms_full, dt, n = cpu_reference_sample(13)
        if i >= args.warmup:
            vals.append(ms_full)
    v = statistics.median(vals)
    sample = "evaluate_in_s restated (per-term scalar mul + into_affine + affine fold) on 2^13 terms per step, scaled x2^11 to 2^24 (the serial loop of equal-cost terms would take ~80 min at full size)"
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
    aff_ms = []
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
        a1 = eng.last_affine_profile()["add_kernel_round1_ms"]
        if a1 > 0:
            aff_ms.append(a1)

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
        aff_ms.clear()
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

    aff = eng.last_affine_profile()
    extra = {}
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": "g1_msm_2p24_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "e2e_ms": e2e_ms, "gpu_launches": timed_launches, "phases_ms": phases,
                              "affine": aff, "quick": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if rank == 0:
        # ---- the other numbers BASELINE.json names, N = 1 semantics on rank 0's GPU ----
        n20 = 1 << 20
        extra["g1_msm_2p24_srs_table_build_s"] = t_tab
        extra["g1_msm_2p24_adhoc_bases_ms"] = timed_local(torch, lambda: eng.msm_dev(scalars, bases, n), 3, 1)
        extra["g1_msm_2p20_adhoc_bases_ms"] = timed_local(torch, lambda: eng.msm_dev(scalars, bases, n20), 5, 3)
        eng.set_msm_affine(0)
        extra["g1_msm_2p24_xyzz_only_ms"] = timed_local(torch, lambda: eng.msm_dev(scalars, None, n), 3, 1)
        eng.set_msm_affine(-1)
        nt = 1 << LOG_N_NTT
        poly = torch.randint(0, 2**62, (nt * 4,), dtype=torch.int64, device=dev)
        ntt_ms = timed_local(torch, lambda: eng.ntt_dev(poly, LOG_N_NTT), 10, 3)
        extra["fr_ntt_2p24_ms"] = ntt_ms
        extra["fr_intt_2p24_ms"] = timed_local(torch, lambda: eng.ntt_dev(poly, LOG_N_NTT, inverse=True), 5, 2)
        extra["fr_coset_ntt_2p24_ms"] = timed_local(torch, lambda: eng.ntt_dev(poly, LOG_N_NTT, coset=7), 5, 2)
        extra["fr_ntt_2p24_launches"] = eng.last_launches("ntt")
        hp = torch.empty(nt * 4, dtype=torch.int64, pin_memory=True)
        hp.copy_(poly)
        hp_np = hp.numpy().view(np.uint64)
        extra["fr_ntt_2p24_e2e_ms"] = timed_local(torch, lambda: eng.ntt(hp_np, LOG_N_NTT), 3, 1)
        hbm, hbm_src = peaks()
        passes_min = -(-LOG_N_NTT // 12)
        ntt_bytes = 64.0 * nt * passes_min
        # (log N - 1) * N / 2 products of 110 IMAD.WIDE + 11 IMAD (SASS of fp_mul<Fr>); the level of order 2 multiplies by 1
        ntt_imad = (nt // 2) * (LOG_N_NTT - 1) * IMAD_PER_FR_MUL_SASS
        extra["ntt_roofline"] = {
            "bound": "hbm", "achieved": ntt_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": ntt_bytes / (ntt_ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_src,
            "algorithmic_bytes": ntt_bytes, "passes_counted": passes_min, "passes_run": eng.last_launches("ntt"),
            "traffic": ntt_traffic(eng.last_launches("ntt")),
            "integer_bound_note": "the transform is bound by the integer pipe, not HBM: (log2 N - 1) * N/2 Fr products x 121 "
                                  "multiply-adds (no inter-pass products since round 2); fraction of the measured IMAD.WIDE "
                                  "peak below.  0.50 of the HBM roofline would need 0.66 ms: out of reach for 255-bit Montgomery "
                                  "arithmetic on this pipe (floor = 2.4 ms at 100 % of the measured multiply-add rate)",
            "int_pipe_frac": ntt_imad / (ntt_ms * 1e-3) / imad_wide if imad_wide else None,
            "int_pipe_floor_ms": ntt_imad / imad_wide * 1e3 if imad_wide else None,
        }
        del poly, hp
        # ---- SRS of 2^20 + 3 points: the 2^20 commitment and the end-to-end PLONK prove (config 4) ----
        extra.update(plonk_and_2p20(z, eng, torch, scalars))

    # ---- multi-GPU rows of SURVEY.md 8(e), measured on ALL ranks (N > 1); at N = 1 the 2^26 MSM on one GPU ----
    eng.srs_upload_dev(bases, 1)  # drops the 2^24 table (18 GiB)
    del bases, host_scalars
    torch.cuda.empty_cache()
    multi = multi_gpu_rows(z, eng, torch, dist, dev, rank, world, scalars)
    if rank == 0:
        extra["multi_gpu"] = multi

    if rank != 0:
        if world > 1:
            dist.barrier()  # rank 0 finishes the host-CPU baselines first
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, integer pipe ----
    # With batched-affine rounds the longest launch is the first round's addition kernel (aff_add_kernel<true>): 5 Fq
    # products per pair-addition (lambda, lambda^2, lambda * dx and the two back-substitution products of the shared
    # inversion); timed live by CUDA events around that launch.  Without rounds it is msm_accumulate_kernel, 10 products
    # per mixed addition.  SURVEY.md 8(d) counts 300 multiply-adds per Fq product.
    phase_imad = float(IMAD_PER_FQ_MUL * FQ_MUL_PER_MADD) * n * n_win  # the survey's unit: 3000 IMAD per (point, window) term
    if aff["rounds"] and aff["add_kernel_round1_ms"] > 0 and len(aff["points"]) >= 2:
        pair_adds = aff["points"][0] - aff["points"][1]
        k_name, k_ms = "aff_add_kernel<first round>", statistics.mean(aff_ms[-args.steps:]) if aff_ms else aff["add_kernel_round1_ms"]
        alg_imad = float(IMAD_PER_FQ_MUL * FQ_MUL_PER_AFFINE_ADD_KERNEL) * pair_adds
        traffic_file = "aff_add_traffic.json"
    else:
        k_name, k_ms, alg_imad, traffic_file = "msm_accumulate_kernel", acc_ms, phase_imad, "msm_accumulate_traffic.json"
    achieved = alg_imad / (k_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", traffic_file)
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {
        "bound": "int-pipe", "kernel": k_name, "achieved": achieved, "peak": imad_wide / 1e12,
        "unit": "TIMAD/s", "frac": achieved / (imad_wide / 1e12) if imad_wide else None,
        "peak_source": "measured in this run: IMAD.WIDE.U32.X carry chains on all SMs (zkp_bench_imad_peak)",
        "peak_imad_lo": imad_lo / 1e12, "frac_of_imad_lo_peak": achieved / (imad_lo / 1e12) if imad_lo else None,
        "algorithmic_imad_per_launch": alg_imad, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms,
        "window_bits": c_bits, "windows": n_win, "traffic": traffic, "phases_ms": phases,
        "affine_rounds": aff["rounds"], "points_per_round": aff["points"],
        "accumulate_phase": {
            "note": "whole bucket-accumulation phase (affine rounds + XYZZ finish) in the survey's units: 3000 multiply-adds per "
                    "(point, window) term; above 1.0 means fewer products were spent than the XYZZ formula needs",
            "ms": acc_ms, "share_of_step": acc_ms / ms, "survey_imad": phase_imad,
            "frac_survey_units": phase_imad / (acc_ms * 1e-3) / imad_wide if imad_wide else None},
    }
    ref_ms, ref_dt, ref_n = cpu_reference_sample(12)
    cpu = {"value": ref_ms, "unit": "ms", "cores": 1, "kind": "port",
           "sample": "evaluate_in_s restated (oracle/zkp_oracle.c:orc_msm_naive) on 2^12 terms = %.2f s, scaled x2^12 to "
                     "2^24; single-threaded like the reference" % ref_dt,
           "best_effort_all_cores": cpu_best_effort(20)}
    try:
        extra["plonk_cpu"] = cpu_plonk_baseline(z, eng)
    except Exception as ex:  # the CPU leg must never take the GPU line down
        extra["plonk_cpu"] = {"error": repr(ex)[:300]}
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
        dist.barrier()
        dist.destroy_process_group()


def multi_gpu_rows(z, eng, torch, dist, dev, rank, world, scratch_scalars):
    """SURVEY.md 8(e) on all `world` ranks.  Every workload is the SAME global problem whatever N (rank-independent
    seeds), so results can be compared across the 1 / 2 / 4 / 8-GPU runs of the scaling sweep."""
    import hashlib

    import numpy as np

    out = {}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- (config 5) strong scaling: ONE 2^26-point MSM, point-range sharded ----
    log_total = 26
    total = 1 << log_total
    lo, hi = z.dist.shard_range(total, rank, world)
    m = hi - lo
    del scratch_scalars
    torch.cuda.empty_cache()
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x2626)
    all_scalars = torch.randint(0, 2**62, (total * 4,), dtype=torch.int64, device=dev, generator=gen)
    sc = all_scalars if world == 1 else all_scalars[lo * 4:hi * 4].clone()
    del all_scalars
    bases = torch.zeros(m * 12, dtype=torch.int64, device=dev)
    eng.generate_bases_dev(0x2627, m, bases, first=lo)
    eng.srs_upload_dev(bases, m)
    del bases
    torch.cuda.empty_cache()
    eng.srs_precompute()
    res = {}

    def msm_step():
        if world == 1:
            res["p"] = eng.msm_dev(sc, None, m)[0]
        else:
            res["p"] = z.dist.msm_sharded(eng, sc, None, m, device=dev)[0]

    ms26 = timed(msm_step, 3, 2)
    c26, w26 = eng.last_msm_shape()
    out["g1_msm_2p26_strong"] = {"ms": ms26, "n_gpus": world, "points_per_gpu": m, "window_bits": c26, "windows": w26,
                                 "affine_rounds": eng.last_affine_rounds(), "mpoints_per_s": total / ms26 / 1e3,
                                 "result_sha256": hashlib.sha256(np.ascontiguousarray(res["p"]).tobytes()).hexdigest(),
                                 "note": "same 2^26 points and scalars at every N (rank-independent seeds): result_sha256 must "
                                         "not change with N; speed-up = this value at N = 1 / this value"}
    eng.srs_upload_dev(sc, 1)
    del sc
    torch.cuda.empty_cache()
    if world == 1:
        return out

    # ---- four-step NTT of 2^26 elements over the ranks, both transports, bit-exact vs the single-GPU transform ----
    log_n = 26
    nn = 1 << log_n
    gen.manual_seed(0x2628)
    x = torch.randint(0, 2**62, (nn * 4,), dtype=torch.int64, device=dev, generator=gen)
    want = x.clone()
    eng.ntt_dev(want, log_n)
    single_ms = timed(lambda: eng.ntt_dev(want, log_n), 3, 1)  # (transforms the buffer repeatedly: timing only)
    want = x.clone()
    eng.ntt_dev(want, log_n)
    row = {"log_n": log_n, "n_gpus": world, "single_gpu_ms": single_ms}
    for p2p in (True, False):
        tag = "fused_peer_exchange" if p2p else "nccl_all_to_all"
        try:
            d = z.dist.DistNtt(eng, log_n, rank, world, p2p=p2p)
            cols = (1 << d.s) // world
            a0 = x.view(1 << d.r, 1 << d.s, 4)[:, rank * cols:(rank + 1) * cols, :].contiguous().view(-1)
            wb = want.view(1 << d.s, 1 << d.r, 4)[:, rank * d.rows_local:(rank + 1) * d.rows_local, :]
            wb = wb.permute(1, 0, 2).contiguous().view(-1)
            b = d.forward(a0.clone())
            ok_f = bool(torch.equal(b, wb))
            ok_i = bool(torch.equal(d.inverse(b.clone()), a0))
            oks = torch.tensor([int(ok_f), int(ok_i)], device=dev)
            dist.all_reduce(oks, op=dist.ReduceOp.MIN)
            work = a0.clone()
            if p2p:
                fwd = timed(lambda: d.forward_into_exchange(work), 5, 2)
            else:
                fwd = timed(lambda: d.forward(work), 5, 2)
            row[tag] = {"forward_ms": fwd, "bit_exact_forward": bool(oks[0].item()), "bit_exact_inverse": bool(oks[1].item()),
                        "speedup_vs_single_gpu": single_ms / fwd}
            d.close()
            del a0, wb, b, work
        except Exception as ex:
            row[tag] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()
    out["fr_ntt_2p26_four_step"] = row
    del x, want
    torch.cuda.empty_cache()

    # ---- batch of 32 independent 2^22 polynomials, sharded whole (no collective on the data path) ----
    batch, lg = 32, 22
    mine = list(z.dist.batch_shard(batch, rank, world))
    data = torch.randint(0, 2**62, (max(len(mine), 1) * (1 << lg) * 4,), dtype=torch.int64, device=dev)
    bms = timed(lambda: eng.ntt_dev(data, lg, batch=len(mine)) if mine else None, 5, 3)
    out["fr_ntt_batch_32x2p22"] = {"ms": bms, "n_gpus": world, "polys_per_gpu": len(mine),
                                   "gelem_per_s": batch * (1 << lg) / bms / 1e6}
    del data
    torch.cuda.empty_cache()

    # ---- sharded PLONK prove, 2^20 gates: replicated prover, point-range-sharded commitments ----
    k = 20
    n = 1 << k
    tot = n + 3
    slo, shi = z.dist.shard_range(tot, rank, world)
    eng.srs_generate(PLONK_SECRET, shi - slo, want_points=False, first=slo)
    eng.srs_precompute()
    cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
    runs, digest = [], None
    for _ in range(4):
        sync()
        t0 = time.perf_counter()
        p = z.plonk.generate_proof_sharded(cc, plonk_blinding(z), rank, world, slo, tot, device=dev)
        runs.append((time.perf_counter() - t0) * 1e3)
        digest = hashlib.sha256(p.to_bytes()).hexdigest()
    best = torch.tensor([min(runs[1:])], device=dev)
    dist.all_reduce(best, op=dist.ReduceOp.MAX)
    digests = [None] * world
    dist.all_gather_object(digests, digest)
    cc.close()
    out["plonk_prove_2p20_sharded"] = {"prove_ms": float(best.item()), "n_gpus": world, "inside_ms": p.timings_ms,
                                       "proof_sha256": digest, "all_ranks_same_bytes": len(set(digests)) == 1}
    return out


def plonk_and_2p20(z, eng, torch, scalars):
    """BASELINE.json config 4 on this GPU: synthetic chain circuits of 2^16 - 3 and 2^20 - 3 gates, SRS from a fixed secret
    with its window table, fixed blinding; wall clock around zkp_plonk_prove (all rounds, transcript included)."""
    import hashlib

    blind = plonk_blinding(z)
    out = {}
    for k in (16, 20):
        n = 1 << k
        eng.srs_generate(PLONK_SECRET, n + 3, want_points=False)
        eng.srs_precompute()
        out["g1_msm_2p%d_ms" % k] = timed_local(torch, lambda: eng.msm_dev(scalars, None, n), 5, 3)
        cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
        runs, wall, digest, inside = [], [], None, None
        for _ in range(4):
            p = z.plonk.generate_proof(cc, blind)
            runs.append(p.timings_ms["total"])
            inside = p.timings_ms
            digest = hashlib.sha256(p.to_bytes()).hexdigest()
        for _ in range(3):  # no timing buffer: the prover never synchronises at phase boundaries
            t0 = time.perf_counter()
            z.plonk.generate_proof(cc, blind, timings=False)
            wall.append((time.perf_counter() - t0) * 1e3)
        cc.close()
        out["plonk_prove_2p%d_ms" % k] = min(min(runs[1:]), min(wall))
        out["plonk_prove_2p%d_split_ms" % k] = inside
        out["plonk_prove_2p%d_runs_ms" % k] = runs + wall
        out["plonk_proof_2p%d_sha256" % k] = digest
    out["plonk_config"] = "chain circuit, 2^k - 3 gates (alternating mul / add, c_i wired to a_(i+1)), seed k, fixed b1..b9"
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
    ap.add_argument("--quick", action="store_true",
                    help="headline step only (no `extra`, no CPU legs): the command profiled under ncu for profiles/")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
