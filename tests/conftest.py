import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def reference_dir():
    """The UNMODIFIED reference tree: $PIC_REFERENCE, /root/reference (build container), or the copy that
    tools/stage_reference.py puts under baseline/_ref/ (git-ignored, but it travels to the GPU box).  None if absent."""
    for p in (os.environ.get("PIC_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.exists(os.path.join(p, "src", "env", "pic.py")):
            return p
    return None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu tests are skipped (not failed) when no device is visible."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]

    return load
