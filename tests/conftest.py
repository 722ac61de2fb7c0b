import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference mounted (build container only)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    import ref_harness
    have_ref = ref_harness.reference_available()
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not mounted"))


@pytest.fixture(scope="session")
def hica():
    """A libhica handle on cuda:0.  No fallback: fails loudly when the library or the GPU is missing."""
    from hic_genome_assembler_b200._lib import Handle
    h = Handle(0)
    yield h
    h.close()
