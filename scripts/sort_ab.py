"""A/B of the MSM grouping step: the engine's own radix sort / scan (csrc/sort.cu, the product) against a build with
cub::DeviceRadixSort / cub::DeviceScan swapped in (-DZKP_USE_CUB).  Build the cub variant first, HERE:
    python scripts/sort_ab.py build
then on the GPU box: python scripts/sort_ab.py"""
import importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CUB = os.path.join(ROOT, "zkp-implementation_b200", "libzkp_b200_cub.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    print(b.build_cuda(force=True, extra_flags=["-DZKP_USE_CUB"], out_name="libzkp_b200_cub.so"))
    sys.exit(0)
import torch
import zkp_implementation_b200 as z

for name, path in (("own", None), ("cub", CUB)):
    if path and not os.path.exists(path):
        continue
    eng = z.Engine(0, lib_path=path)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_profiling(True)
    for log_n in (20, 24):
        n = 1 << log_n
        bases = torch.zeros(n * 12, dtype=torch.int64, device="cuda")
        eng.generate_bases_dev(0xB200, n, bases)
        s = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
        eng.srs_upload_dev(bases, n)
        del bases
        eng.srs_precompute()
        for _ in range(3):
            out = eng.msm_dev(s, None, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = eng.msm_dev(s, None, n)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"sort": name, "log_n": log_n, "ms": e0.elapsed_time(e1) / 5, "phases_ms": eng.last_phase_ms(),
                          "x0": hex(int(out[0][0]))}), flush=True)
        eng.srs_upload_dev(s, 1)
        del s
        torch.cuda.empty_cache()
    eng.close()
