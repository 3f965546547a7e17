"""ncu target: fixed-base MSM over a resident SRS of 2^PROF_LOG_N points (one warm + one measured call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkp_implementation_b200 as z

log_n = int(os.environ.get("PROF_LOG_N", "18"))
n = 1 << log_n
eng = z.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
eng.generate_bases_dev(0xB200, n, bases)
s = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
eng.srs_upload_dev(bases, n)
eng.srs_precompute(int(os.environ.get("PROF_BITS", "0")))
torch.cuda.synchronize()
print("MARK", flush=True)
for _ in range(2):
    eng.msm_dev(s, None, n)
torch.cuda.synchronize()
print("ok", eng.last_msm_shape(), eng.last_phase_ms())
