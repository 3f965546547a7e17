"""Quick NTT timing + parity probe (development aid, not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkp_implementation_b200 as z
from oracle import coracle as c
eng = z.Engine(0, lib_path=os.environ.get("ZKP_LIB")); eng.set_stream(torch.cuda.current_stream().cuda_stream)
F = z.fields
a = F.random_fr_mont(1, 1 << 20)
t = torch.from_numpy(a.view(np.int64)).cuda()
eng.ntt_dev(t, 20)
print("parity 2^20:", bool((t.cpu().numpy().view(np.uint64).reshape(-1, 4) == c.ntt(a, 20)).all()))
for log_n, batch in ((16, 1), (20, 1), (22, 1), (24, 1), (26, 1), (20, 16), (22, 8)):
    n = 1 << log_n
    x = torch.randint(0, 2**62, (n * 4 * batch,), dtype=torch.int64, device="cuda")
    for _ in range(3): eng.ntt_dev(x, log_n, batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng.ntt_dev(x, log_n, batch)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("ntt 2^%d x%d: %.3f ms  (%.2f ns/elem)" % (log_n, batch, ms, ms * 1e6 / (n * batch)))
    del x
