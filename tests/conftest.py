import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def cuda_ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from meta_interpolation_b200 import backbone
    ops = backbone.default_ops()   # raises if libmi_b200.so is missing: no silent fallback
    # Dirty the caching allocator once: a 1 GiB block of NaN is released back to it, so the `torch.empty` buffers of
    # the tests that follow are carved out of NaN-filled memory.  A kernel that reads what nobody wrote (or writes
    # into a neighbouring channel slice) then fails here too, not only under MI_B200_POISON=1 or after some unlucky
    # sequence of earlier tests (see profiles/README.md for the bug this caught).
    dirty = torch.full((256 * 1024 * 1024,), float("nan"), device="cuda")
    del dirty
    return ops


@pytest.fixture()
def ref_ops():
    from oracle.ops_ref import RefOps
    from meta_interpolation_b200 import backbone
    ops = RefOps()
    saved = backbone._default_ops
    backbone.set_default_ops(ops)
    yield ops
    backbone.set_default_ops(saved)
