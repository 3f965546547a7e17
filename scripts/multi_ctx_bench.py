"""Single-process multi-GPU context (zkp_ctx_create_multi): one 2^log_n-point commitment against the sharded resident
SRS, host scalars (zkp_msm_g1) and device-resident scalars (zkp_msm_g1_dev), over 1..G devices of this box in ONE
process.  Wall clock around the call (the per-device streams are synchronised inside it); the result must be the same
point for every G.   usage: multi_ctx_bench.py [log_n] [max_devices]"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
gmax = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
n = 1 << log_n
host = torch.empty(n * 4, dtype=torch.int64, pin_memory=True)
gen = torch.Generator(device="cuda:0")
gen.manual_seed(0x77)
host.copy_(torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda:0", generator=gen))
host_np = host.numpy().view(np.uint64)
g = 1
while g <= gmax:
    eng = z.Engine(devices=list(range(g)))
    t0 = time.perf_counter()
    eng.srs_generate(0xB200B200, n, want_points=False)
    eng.srs_precompute()
    setup = time.perf_counter() - t0
    dev = torch.empty(n * 4, dtype=torch.int64, device="cuda:0")
    dev.copy_(host)
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("host_scalars", lambda: eng.msm(host_np)), ("device_scalars", lambda: eng.msm_dev(dev, None, n))):
        for _ in range(2):
            out = fn()[0]
        t0 = time.perf_counter()
        for _ in range(3):
            out = fn()[0]
        res[name + "_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        res[name + "_sha256"] = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()[:16]
    print(json.dumps({"op": "multi_ctx_msm", "log_n": log_n, "devices": g, "setup_s": setup, **res}), flush=True)
    eng.close()
    del dev
    torch.cuda.empty_cache()
    g *= 2
