import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `-m gpu`)")
    # the shared libraries are build artefacts (git-ignored): build them once if a fresh checkout lacks them
    lib = os.path.join(ROOT, "scoary_b200", "libscoary_b200.so")
    ora = os.path.join(ROOT, "oracle", "libscoary_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__
        __graft_entry__.build()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    """One engine context per session.  GPU tests must not pass on a fallback:
    Engine() raises when libscoary_b200.so is missing or no B200 is present."""
    from scoary_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def pytest_collection_modifyitems(config, items):
    # the newest GPU checks and the full-size runs (BASELINE.json's shapes, seconds each) go last: `-x` reaches
    # them only after every earlier parity test has passed
    items.sort(key=lambda it: 2 if "test_gpu_full_size" in it.nodeid else 1 if "test_gpu_cli_more" in it.nodeid else 0)
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
