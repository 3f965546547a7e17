import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _build_mod():
    import importlib.util

    spec = importlib.util.spec_from_file_location("zkp_b200_build", os.path.join(ROOT, "zkp-implementation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def zkp():
    import zkp_implementation_b200 as z

    return z


@pytest.fixture(scope="session")
def coracle():
    from oracle import coracle as c

    c.build()
    return c


@pytest.fixture(scope="session")
def pyref():
    from oracle import pyref

    return pyref


@pytest.fixture(scope="session")
def hostlib():
    import ctypes

    return ctypes.CDLL(_build_mod().build_hosttest())


@pytest.fixture(scope="session")
def emu_engine(zkp):
    """Engine bound to the CPU kernel emulator (tests/emu): checks kernel indexing logic without a GPU."""
    path = _build_mod().build_emu()
    eng = zkp.Engine(0, lib_path=path)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def gpu_engine(zkp):
    """Engine on cuda:0 through the real library; never falls back to anything else."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    eng = zkp.Engine(0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    yield eng
    eng.close()


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def engine(request):
    """Same test body on the CPU emulator (kernel logic) and on the B200 (the product path)."""
    return request.getfixturevalue("emu_engine" if request.param == "emu" else "gpu_engine")
