import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _built(path, target):
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(path)


@pytest.fixture(scope="session")
def port():
    from oracle import orc
    assert _built(orc.PATHS["port"], "port"), "oracle/liboracle.so could not be built"
    return orc


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libfovref.so: the reference's sources compiled unmodified.  Present in the build
    container (and shipped prebuilt to the GPU box); elsewhere tests fall back to tests/golden."""
    from oracle import orc
    if not os.path.exists(orc.PATHS["reference"]):
        if os.path.isdir("/root/reference/PT_sv5_"):
            _built(orc.PATHS["reference"], "ref")
    if not os.path.exists(orc.PATHS["reference"]):
        pytest.skip("oracle/_ref not built and /root/reference absent (golden fixtures cover this)")
    return orc


GOLDEN = os.path.join(ROOT, "tests", "golden")
