"""Development aid: time zkp_sort_pairs_dev (201M pairs, 22-bit keys = the 2^24 MSM's grouping step) for build variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zkp_implementation_b200 as z
n = 201326592
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys0 = torch.randint(0, 1 << 21, (n,), dtype=torch.int32, device="cuda", generator=g)
vals0 = torch.arange(n, dtype=torch.int32, device="cuda")
for name in ("b200",):
    path = os.path.join(ROOT, "zkp-implementation_b200", "libzkp_%s.so" % name)
    if not os.path.exists(path):
        continue
    eng = z.Engine(0, lib_path=path); eng.set_stream(torch.cuda.current_stream().cuda_stream)
    k, v = keys0.clone(), vals0.clone()
    eng.sort_pairs_dev(k, v, n, 22)
    torch.cuda.synchronize()
    ok = bool((k[1:] >= k[:-1]).all())
    ts = []
    for _ in range(4):
        k.copy_(keys0); v.copy_(vals0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.sort_pairs_dev(k, v, n, 22); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "sorted" if ok else "NOT SORTED", "ms: %.3f" % min(ts), flush=True)
    eng.close()
