"""Development aid: time NTT build variants (compile-time knobs) against each other on the GPU box.
The variant libraries are built HERE (nvcc cross-compiles) before the gpurun call: python scripts/ntt_variants.py build"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "base": [],
    "k1_b7": ["-DNTT_KMAX=1", "-DNTT_MIN_BLOCKS=7"],
    "k1_b6": ["-DNTT_KMAX=1", "-DNTT_MIN_BLOCKS=6"],
    "k2_b6": ["-DNTT_MIN_BLOCKS=6"],
    "k2_th256_b3_t11": ["-DNTT_TILE_LOG=11", "-DNTT_THREADS_PER_CTA=256", "-DNTT_MIN_BLOCKS=2"],
    "k2_th64_b10": ["-DNTT_THREADS_PER_CTA=64", "-DNTT_MIN_BLOCKS=10"],
}
if len(sys.argv) > 1 and sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    for name, flags in VARIANTS.items():
        print(name, b.build_cuda(force=True, extra_flags=flags, out_name="libzkp_var_%s.so" % name, only=["ntt.cu"]))
    sys.exit(0)
import numpy as np, torch
import zkp_implementation_b200 as z
from oracle import coracle as c
F = z.fields
a = F.random_fr_mont(1, 1 << 20)
want = c.ntt(a, 20)
for name in VARIANTS:
    path = os.path.join(ROOT, "zkp-implementation_b200", "libzkp_var_%s.so" % name)
    if not os.path.exists(path):
        continue
    eng = z.Engine(0, lib_path=path); eng.set_stream(torch.cuda.current_stream().cuda_stream)
    t = torch.from_numpy(a.view(np.int64).copy()).cuda()
    eng.ntt_dev(t, 20)
    ok = bool((t.cpu().numpy().view(np.uint64).reshape(-1, 4) == want).all())
    row = [name, "ok" if ok else "MISMATCH"]
    for log_n in (20, 22, 24):
        n = 1 << log_n
        x = torch.randint(0, 2**62, (n * 4,), dtype=torch.int64, device="cuda")
        for _ in range(3): eng.ntt_dev(x, log_n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): eng.ntt_dev(x, log_n)
        e1.record(); torch.cuda.synchronize()
        row.append("2^%d %.3f ms" % (log_n, e0.elapsed_time(e1) / 8))
        del x
    print(*row, flush=True)
    eng.close()
