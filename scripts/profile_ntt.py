"""ncu target: forward Fr NTT of 2^PROF_LOG_N elements resident in HBM (two warm-ups + one measured call = 3 passes each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkp_implementation_b200 as z

log_n = int(os.environ.get("PROF_LOG_N", "24"))
eng = z.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
p = torch.randint(0, 2**62, ((1 << log_n) * 4,), dtype=torch.int64, device="cuda")
for _ in range(3):
    eng.ntt_dev(p, log_n)
torch.cuda.synchronize()
print("ok", eng.last_launches("ntt"))
