import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def hostcheck():
    """ctypes handle on the TEST-ONLY host emulation of the kernel bodies (tests/hostcheck/hostcheck.cpp)."""
    import ctypes

    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    so = os.path.join(ROOT, "tests", "hostcheck", "libefb_hostcheck.so")
    deps = [src] + [os.path.join(ROOT, "easyfea_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "easyfea_b200", "csrc"))
                    if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)
