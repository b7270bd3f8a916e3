import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if lib.cuInit(0) != 0:
            return False
        lib.cuDeviceGetCount(ctypes.byref(n))
        return n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run under gpurun / at round end)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_port():
    """Build (if needed) and return the name of the plain-C oracle."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return "port"


@pytest.fixture(scope="session")
def oracle_ref():
    """The reference's own code (oracle/_ref). Built here when /root/reference exists; on
    the GPU box only the prebuilt library is available."""
    from oracle import build_ref, orc
    if build_ref.available():
        build_ref.build(verbose=False)
    if not orc.have("reference"):
        pytest.skip("oracle/_ref not built and no reference tree to build it from")
    return "reference"
