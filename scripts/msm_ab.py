"""A/B of the bucket accumulation: XYZZ only (affine rounds = 0) vs batched-affine tree rounds, fixed-base 2^log_n MSM.
usage: msm_ab.py [log_n] [rounds ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkp_implementation_b200 as z  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rounds_list = [int(x) for x in sys.argv[2:]] or [0, -1]
n = 1 << log_n
if os.environ.get("ZKP_L2_FETCH"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    print("cudaDeviceSetLimit(MaxL2FetchGranularity) ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["ZKP_L2_FETCH"]))))
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("granularity now", v.value)
eng = z.Engine(0, lib_path=os.environ.get("ZKP_LIB"))
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_profiling(os.environ.get("ZKP_PROFILING", "1") != "0")  # profiling on = per-phase events, affine rounds on ONE stream
bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
eng.generate_bases_dev(0xB200, n, bases)
scalars = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
eng.srs_upload_dev(bases, n)
eng.srs_precompute(int(os.environ.get("ZKP_C", "0")))
ref = None
for r in rounds_list:
    eng.set_msm_affine(r)
    for _ in range(2):
        out = eng.msm_dev(scalars, None, n)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 4
    for _ in range(steps):
        eng.msm_dev(scalars, None, n)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    print(json.dumps({"log_n": log_n, "affine_rounds_req": r, "affine_rounds": eng.last_affine_rounds(),
                      "ms": e0.elapsed_time(e1) / steps, "same_as_first": bool((out == ref).all()),
                      "phases_ms": eng.last_phase_ms(), "launches": eng.last_launches("msm"),
                      "shape": eng.last_msm_shape()}), flush=True)
