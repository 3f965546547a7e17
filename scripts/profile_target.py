"""Small fixed workload for ncu captures: one warm + one measured MSM 2^24, NTT 2^24 fwd (3 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkp_implementation_b200 as z

log_n = int(os.environ.get("PROF_LOG_N", "24"))
n = 1 << log_n
eng = z.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
eng.generate_bases_dev(0xB200, n, bases)
s = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.msm_dev(s, bases, n)
p = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.ntt_dev(p, log_n)
torch.cuda.synchronize()
print("ok")
