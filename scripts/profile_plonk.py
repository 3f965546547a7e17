"""ncu target: one warm + one measured PLONK prove of a 2^PROF_LOG_N chain circuit (device-resident prover)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z

k = int(os.environ.get("PROF_LOG_N", "16"))
n = 1 << k
eng = z.Engine(0)
eng.srs_generate(0x1F2E3D4C5B6A79881234567, n + 3, want_points=False)
eng.srs_precompute()
cc = z.plonk.chain_circuit(n - 3, seed=k).compile(eng)
blind = [(0xABCDEF0123456789 * (i + 3) ** 7) % z.FR_MODULUS for i in range(9)]
for _ in range(2):
    p = z.plonk.generate_proof(cc, blind)
print("MARK prove_ms", p.timings_ms)
cc.close()
