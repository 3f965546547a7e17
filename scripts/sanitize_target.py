"""compute-sanitizer target: the smoke path plus the sort / polynomial layer at small sizes (no torch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
import zkp_implementation_b200 as z

g.smoke()
eng = z.Engine(0)
rng = np.random.default_rng(1)
n = 50000
k = rng.integers(0, 1 << 21, size=n, dtype=np.uint64).astype(np.uint32)
v = np.arange(n, dtype=np.uint32)
dk, dv = eng.dev_alloc(n * 4), eng.dev_alloc(n * 4)
eng._check(eng.lib.zkp_dev_upload(eng._h, dk, k.ctypes.data, n * 4))
eng._check(eng.lib.zkp_dev_upload(eng._h, dv, v.ctypes.data, n * 4))
eng.sort_pairs_dev(dk, dv, n, 21)
out = np.zeros(n, dtype=np.uint32)
eng._check(eng.lib.zkp_dev_download(eng._h, out.ctypes.data, dk, n * 4))
assert (out == np.sort(k, kind="stable")).all()
F = z.fields
# fixed-base MSM + batch at a size with several reduction levels
srs = z.Srs.new_from_secret(eng, 77, 5000)
z.KzgScheme(eng, srs)
s = F.random_fr_mont(5, 5003)
print("msm", eng.msm(s)[1], "launches", eng.last_launches("msm"))
a = eng.vec(F.random_fr_mont(6, 9000)); eng.fr_scan(a, "mul"); eng.fr_batch_inverse(a)
print("sanitize target OK")
# round 2: batched-affine rounds (forced), windowed and fixed-base, odd sizes; NTT with the per-level tables
eng.set_msm_affine(3)
s2 = F.random_fr_mont(7, 5003)
pts = srs.g1_limbs()
print("msm affine windowed", eng.msm(s2, pts)[1], "rounds", eng.last_affine_rounds())
eng.srs_precompute(7)
print("msm affine fixed", eng.msm(s2)[1], "rounds", eng.last_affine_rounds())
eng.set_msm_affine(-1)
d = F.random_fr_mont(8, 1 << 13)
eng.ntt(d, 13)
eng.ntt(d, 13, inverse=True, coset=7)
print("sanitize target round-2 OK")
