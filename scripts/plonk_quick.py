"""PLONK prove timing on one GPU for a few sizes (development aid): prints total / msm / ntt / other and the proof hash."""
import hashlib, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z
SECRET = 0x1F2E3D4C5B6A79881234567
BLIND = [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]
eng = z.Engine(0, lib_path=os.environ.get("ZKP_LIB"))
for k in [int(a) for a in sys.argv[1:]] or [16, 20]:
    n = 1 << k
    eng.srs_generate(SECRET, n + 3, want_points=False)
    eng.srs_precompute()
    sc = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
    for _ in range(3): eng.msm_dev(sc, None, n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): eng.msm_dev(sc, None, n)
    msm_ms = (time.perf_counter() - t0) / 5 * 1e3
    cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
    runs = []
    for _ in range(4):
        p = z.plonk.generate_proof(cc, BLIND)
        runs.append(p.timings_ms)
    print(json.dumps({"log_n": k, "msm_ms": msm_ms, "shape": eng.last_msm_shape(), "prove": runs[-1],
                      "sha": hashlib.sha256(p.to_bytes()).hexdigest()[:16]}), flush=True)
    cc.close()
