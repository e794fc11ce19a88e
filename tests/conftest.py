import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """CPU model of the provisional FLP0 bitstream (NOT the reference; see oracle/flp0_oracle.h)."""
    so = os.path.join(ROOT, "oracle", "libflp0_oracle.so")
    src = os.path.join(ROOT, "oracle", "flp0_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    import oracle_binding

    return oracle_binding.Oracle(so)


@pytest.fixture(scope="session")
def flic():
    import flic_b200

    flic_b200.build_library()
    return flic_b200


@pytest.fixture(scope="session")
def codec(flic):
    """Engine context on cuda:0. Fails (does not skip) if the CUDA library cannot run: no CPU fallback."""
    c = flic.Codec(0)
    yield c
    c.close()
