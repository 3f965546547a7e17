"""One 2^log_n forward NTT repeated a few times (profiling target: `ncu -k regex:ntt_pass`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = z.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
data = torch.randint(0, 2**62, ((1 << log_n) * 4,), dtype=torch.int64, device="cuda")
eng.ntt_dev(data, log_n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.ntt_dev(data, log_n)
e1.record()
torch.cuda.synchronize()
print("ntt 2^%d: %.3f ms" % (log_n, e0.elapsed_time(e1) / reps))
