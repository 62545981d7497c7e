import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build the native pieces once if they are missing (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "needle_b200", "libneedle_b200.so")
    ora = os.path.join(ROOT, "oracle", "libneedle_oracle.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "needle_b200", "csrc")], check=True)
    if not os.path.exists(ora):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
