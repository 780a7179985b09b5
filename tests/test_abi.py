"""The C-ABI library loads and exports exactly what include/mi_b200.h declares (no GPU needed)."""
import ctypes as C
import os

from abi_util import parse_header

from meta_interpolation_b200 import _lib


def _ctype(param):
    if "*" in param:
        return C.c_void_p
    base = param.rsplit(" ", 1)[0]
    return {"int": C.c_int, "float": C.c_float, "size_t": C.c_size_t, "mi_stream_t": C.c_void_p}[base]


def test_library_is_built_in_tree():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    header = parse_header()
    lib = _lib.load()
    assert len(header) >= 33
    for name, (ret, params) in header.items():
        assert hasattr(lib, name), "libmi_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes binding for %s" % name
        assert [_ctype(p) for p in params] == _lib.SIGNATURES[name][1], name
    assert set(_lib.SIGNATURES) == set(header)


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.mi_version() >= 100
    assert b"workspace" in lib.mi_error_string(10003)
    assert lib.mi_launch_count() == 0 or lib.mi_launch_count() > 0


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_cuda_ops_refuses_without_gpu():
    import pytest
    import torch
    from meta_interpolation_b200.ops import CudaOps
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MiB200Error):
        CudaOps()
