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
