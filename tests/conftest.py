import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The shared library is a build artefact (git-ignored).  A fresh checkout builds it once here, exactly as
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU); a failed build fails the session.
    so = os.path.join(ROOT, "munit_b200", "libmunit_b200.so")
    if not os.path.isfile(so):
        import subprocess

        subprocess.check_call(["bash", os.path.join(ROOT, "munit_b200", "csrc", "build.sh")])


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name), weights_only=False)

    return load
