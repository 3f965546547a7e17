"""Build recipes for the engine's native libraries (in-tree, so the .so travels with `gpurun`).

* ``build_cuda()``  nvcc, sm_100a only -> zkp-implementation_b200/libzkp_b200.so  (the product)
* ``build_hosttest()``  g++ -> zkp-implementation_b200/libzkp_hosttest.so (field/curve headers on the
  host, used by the CPU test-suite to check the arithmetic the kernels are built from)
* ``build_emu()``  g++ -DZKP_EMU -> tests/emu/_build/libzkp_b200_emu.so (kernel-logic emulator for the
  CPU test-suite; never loaded by the product path)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "msm_affine.cu", "gen.cu", "poly.cu", "sort.cu"]
HEADERS = ["field.cuh", "curve.cuh", "inv_gcd.cuh", "memops.cuh", "engine.h", "runtime.h", "poly.h"]
HOST_DIR = os.path.join(PKG, "host")
HOST_SOURCES = ["plonk.cpp", "kzg.cpp", "transcript_api.cpp"]  # host orchestration above the C ABI (include/zkp_plonk.h), plain g++
HOST_HEADERS = ["mont_host.hpp", "transcript.hpp"]
HOST_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-fopenmp", "-Wall"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd: list[str], log: str | None = None) -> None:
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def _deps() -> list[str]:
    return ([os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HOST_DIR, f) for f in HOST_SOURCES + HOST_HEADERS]
            + [os.path.join(ROOT, "include", h) for h in ("zkp_b200.h", "zkp_plonk.h")])


def _host_objs(bdir: str) -> list[str]:
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(bdir, src.replace(".cpp", ".host.o"))
        _run(["g++", *HOST_FLAGS, "-c", os.path.join(HOST_DIR, src), "-o", obj])
        objs.append(obj)
    return objs


def build_cuda(force: bool = False, extra_flags: list[str] | None = None, out_name: str = "libzkp_b200.so",
               only: list[str] | None = None) -> str:
    """`only`: build-variant shortcut -- recompile just these sources with `extra_flags` and link them with the
    objects of the default build (which must exist)."""
    out = os.path.join(PKG, out_name)
    if not force and _newer(out, _deps() + [os.path.abspath(__file__)]):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    bdir = os.path.join(PKG, "build", out_name.replace(".so", ""))
    os.makedirs(bdir, exist_ok=True)
    flags = NVCC_FLAGS + (extra_flags or [])

    base_dir = os.path.join(PKG, "build", "libzkp_b200")

    def one(src: str) -> str:
        if only is not None and src not in only:
            return os.path.join(base_dir, src.replace(".cu", ".o"))
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        _run([nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj], log=obj + ".log")
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, SOURCES))
    if only is not None:
        objs += [os.path.join(base_dir, s.replace(".cpp", ".host.o")) for s in HOST_SOURCES]
    else:
        objs += _host_objs(bdir)
    # --cudart shared: the CUDA runtime is resolved from the image (or shared with torch when torch loaded it first)
    # instead of being embedded in the product library
    _run([nvcc, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp",
          "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", out, *objs])
    return out


def build_hosttest(force: bool = False) -> str:
    out = os.path.join(PKG, "libzkp_hosttest.so")
    deps = [os.path.join(CSRC, f) for f in ("host_testapi.cpp", "field.cuh", "curve.cuh", "inv_gcd.cuh")]
    if not force and _newer(out, deps):
        return out
    _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + CSRC, os.path.join(CSRC, "host_testapi.cpp"), "-o", out])
    return out


def build_emu(force: bool = False) -> str:
    edir = os.path.join(ROOT, "tests", "emu")
    out = os.path.join(edir, "_build", "libzkp_b200_emu.so")
    deps = _deps() + [os.path.join(edir, "cuda_emu.h")]
    if not force and _newer(out, deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)

    def one(src: str) -> str:
        obj = os.path.join(edir, "_build", src.replace(".cu", ".o"))
        _run(["g++", "-O2", "-std=c++17", "-DZKP_EMU", "-fPIC", "-pthread", "-I" + edir, "-I" + CSRC, "-x", "c++", "-c",
              os.path.join(CSRC, src), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, SOURCES))
    objs += _host_objs(os.path.join(edir, "_build"))
    _run(["g++", "-shared", "-pthread", "-fopenmp", "-o", out, *objs])
    return out


if __name__ == "__main__":
    what = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cuda", "hosttest"]
    for w in what:
        print(w, "->", {"cuda": build_cuda, "hosttest": build_hosttest, "emu": build_emu}[w](force="--force" in sys.argv))
