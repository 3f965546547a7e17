"""Fixed-base window width sweep: ms per MSM for every c around the cost model's choice.  usage: window_sweep.py log_n [c ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

log_n = int(sys.argv[1])
cs = [int(x) for x in sys.argv[2:]] or [0]
n = 1 << log_n
eng = z.Engine(0, lib_path=os.environ.get("ZKP_LIB"))
eng.set_stream(torch.cuda.current_stream().cuda_stream)
bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
eng.generate_bases_dev(0xB200, n, bases)
scalars = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
eng.srs_upload_dev(bases, n)
ref = None
for c in cs:
    eng.srs_precompute(c)
    for _ in range(3):
        out = eng.msm_dev(scalars, None, n)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 8
    for _ in range(steps):
        eng.msm_dev(scalars, None, n)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    print(json.dumps({"log_n": log_n, "c_req": c, "shape": eng.last_msm_shape(), "rounds": eng.last_affine_rounds(),
                      "ms": round(e0.elapsed_time(e1) / steps, 4), "same": bool((out == ref).all())}), flush=True)
