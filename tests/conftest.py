import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tmpdir_repo():
    d = os.path.join(ROOT, "tests", "_tmp")
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared objects must exist; build them if a fresh checkout has none."""
    import subprocess

    if not os.path.exists(os.path.join(ROOT, "oracle", "libldcore.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libldcore.so"])
    if not os.path.exists(os.path.join(ROOT, "tomahawk_b200", "libtwkb.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tomahawk_b200", "csrc")])
