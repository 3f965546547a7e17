"""End-to-end PLONK prove (BASELINE.json config 4) on one B200: synthetic chain circuit of 2^k - 3 gates,
SRS from a fixed secret generated on the device, fixed blinding.  Prints one JSON line per size with the
prover's own breakdown (MSM calls / NTT + product calls / host arithmetic, wall clock inside zkp_plonk_prove)."""
import hashlib
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

SECRET = 0x1F2E3D4C5B6A79881234567
BLIND = [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]


def main():
    logs = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
    eng = z.Engine(0)
    for k in logs:
        n = 1 << k
        t0 = time.perf_counter()
        eng.srs_generate(SECRET, n + 3, want_points=False)
        t_srs = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.srs_precompute()
        t_tab = time.perf_counter() - t0
        t0 = time.perf_counter()
        circ = z.plonk.chain_circuit(n - 3, seed=k)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        cc = circ.compile(eng)
        t_compile = time.perf_counter() - t0
        runs = []
        digest = None
        for it in range(3):
            p = z.plonk.generate_proof(cc, BLIND)
            runs.append(p.timings_ms)
            d = hashlib.sha256(p.to_bytes()).hexdigest()
            assert digest in (None, d)
            digest = d
        best = min(runs, key=lambda r: r["total"])
        print(json.dumps({"log_n": k, "prove_ms": best["total"], "msm_ms": best["msm"], "ntt_ms": best["ntt"],
                          "other_ms": best["other"], "first_run_ms": runs[0]["total"], "srs_generate_s": t_srs, "srs_precompute_s": t_tab,
                          "circuit_build_s": t_build, "compile_s": t_compile, "proof_sha256": digest}), flush=True)
        cc.close()


if __name__ == "__main__":
    main()
