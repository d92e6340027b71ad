import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """``gpu`` tests need a CUDA device AND the built library: skip them (instead of 100 failures) anywhere else.
    On a GPU box a missing library is a hard error -- the product has no fallback, and a silent skip would hide it."""
    import torch
    lib = os.path.join(ROOT, "iccv19_vqa-cti_b200", "libcti_sm100.so")
    if torch.cuda.is_available():
        if not os.path.exists(lib) and any("gpu" in it.keywords for it in items):
            raise pytest.UsageError(f"{lib} is missing on a GPU box: run `python __graft_entry__.py build` first")
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "cti_golden.pt"), weights_only=False)
