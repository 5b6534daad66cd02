import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must never silently pass on a CPU box
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_built():
    """libhuffb200.so + oracle built (prebuilt files are reused; make is a no-op then)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def orc(native_built):
    import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref(native_built):
    """The unmodified reference CPU path, or None when oracle/_ref/libref.so was not shipped."""
    import pyoracle
    return pyoracle.try_ref()


@pytest.fixture(scope="session")
def hb(native_built):
    import huffman_gpu_b200
    return huffman_gpu_b200


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def c1():
    g = load_golden("c1_fixture.json")
    for k in ("freqs", "codewords", "codewordlens"):
        g[k] = np.array(g[k], dtype=np.uint64 if k == "freqs" else np.uint32)
    return g
