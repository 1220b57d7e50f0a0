"""pytest configuration: markers, import paths, shared fixture loaders."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_sessionstart(session):
    """Make sure every native piece exists before tests import it (no-op when up to date): the
    CUDA library and the marshalling extension are built in-tree and are not part of the git
    history, so a fresh checkout has to compile them once (nvcc cross-compiles without a GPU)."""
    try:
        import __graft_entry__ as G
        G.build_cuda()
        G.build_marshal()
        G.build_oracle()
        G.build_emu()
    except Exception as exc:  # let the individual tests report what is missing
        print("conftest: native build step failed: %r" % (exc,))


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of erroring
    on the engine fixture (the product itself still fails loudly without CUDA: that is a test of its own)."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def ints(xs):
    return [int(x) for x in xs]
