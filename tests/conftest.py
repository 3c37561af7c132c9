import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run on the GPU box with `pytest -m gpu`")


@pytest.fixture(scope="session")
def hostsim():
    """g++ build of the kernels' __host__ __device__ math (tests/hostsim): CPU-side check of the formulas only."""
    import ctypes
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    lib = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    hdr = os.path.join(ROOT, "mli_nerf_b200", "csrc", "mli_math.h")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", lib, src])
    return ctypes.CDLL(lib)


@pytest.fixture(scope="session")
def mli_lib():
    from mli_nerf_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib
