import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def ref_bin():
    """Path of oracle/_ref (the reference's own sources compiled), or None."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "sam2pairs")) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False, capture_output=True)
    return d if os.path.exists(os.path.join(d, "sam2pairs")) else None
