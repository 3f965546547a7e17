"""The C-ABI library loads on a CPU-only box, exports every symbol include/zkp_b200.h declares, and
refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda_lib_path():
    import importlib.util

    spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_cuda()  # nvcc cross-compiles sm_100a without a GPU


def _declared_symbols(header="zkp_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(zkp, cuda_lib_path):
    syms = _declared_symbols()
    assert len(syms) >= 20
    assert sorted(zkp.ABI.keys()) == syms, "python binding table and header disagree"
    lib = zkp.load_library(cuda_lib_path)
    for s in syms:
        assert hasattr(lib, s), s


def test_plonk_header_symbols_all_exported(zkp, cuda_lib_path):
    """include/zkp_plonk.h (host orchestration of the prover) ships in the same library."""
    syms = _declared_symbols("zkp_plonk.h")
    assert sorted(zkp.plonk.PLONK_ABI.keys()) == syms
    lib = zkp.load_library(cuda_lib_path)
    for s in syms:
        assert hasattr(lib, s), s


def test_missing_library_fails_loudly(zkp, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        zkp.load_library(str(tmp_path / "libzkp_b200.so"))


def test_no_device_is_an_error_not_a_fallback(zkp, cuda_lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zkp.ZkpError) as ei:
        zkp.Engine(0, lib_path=cuda_lib_path)
    assert ei.value.status == 6  # ZKP_B200_ERR_NO_DEVICE


def test_strerror(zkp, cuda_lib_path):
    lib = zkp.load_library(cuda_lib_path)
    assert lib.zkp_strerror(0) == b"ok"
    assert b"polynomial.degree()" in lib.zkp_strerror(4)
    assert b"no CPU fallback" in lib.zkp_strerror(6)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "zkp-implementation_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|libzkp_oracle|zkp_oracle\.c|coracle|pyref\.", re.M)
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), (dirpath, f)
