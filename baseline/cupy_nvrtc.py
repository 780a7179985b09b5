"""A stand-in for the two `cupy` entry points the reference's separable-convolution operator uses
(reference sepconv/sepconv_op/sepconv.py:1, 240-243, 276-291): `cupy.util.memoize` and
`cupy.cuda.compile_with_cache(src).get_function(name)(grid=, block=, args=, stream=)`.

cupy is not in this image; NVRTC and the driver API are (cuda-python).  The reference's OWN kernel strings are
compiled for the device at hand and launched on the stream the reference passes, so `FunctionSepconv` runs exactly
the code the authors wrote.  Reference-arm / test infrastructure only: nothing in `meta_interpolation_b200/`
imports this file.
"""
import ctypes
import functools
import sys
import types

from cuda.bindings import driver, nvrtc


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError("CUDA/NVRTC error %s" % (err,))
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


class _Function:
    def __init__(self, module, name):
        self.func = _check(driver.cuModuleGetFunction(module, name.encode()))

    def __call__(self, grid, block, args, stream=None, shared_mem=0):
        # the reference passes python ints: element counts (int32 `const int n`) and raw device addresses
        holders, ptrs = [], []
        for i, a in enumerate(args):
            h = ctypes.c_int(a) if i == 0 else ctypes.c_void_p(a)
            holders.append(h)
            ptrs.append(ctypes.addressof(h))
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        s = getattr(stream, "ptr", 0) if stream is not None else 0
        _check(driver.cuLaunchKernel(self.func, grid[0], grid[1], grid[2], block[0], block[1], block[2], shared_mem,
                                     s, ctypes.addressof(arr), 0))


class _Module:
    def __init__(self, source):
        import torch
        major, minor = torch.cuda.get_device_capability()
        arch = "sm_%d%d%s" % (major, minor, "a" if major >= 9 else "")
        prog = _check(nvrtc.nvrtcCreateProgram(source.encode(), b"kernel.cu", 0, [], []))
        opts = [("--gpu-architecture=" + arch).encode()]
        res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if int(res[0]) != 0:
            size = _check(nvrtc.nvrtcGetProgramLogSize(prog))
            log = b" " * size
            nvrtc.nvrtcGetProgramLog(prog, log)
            raise RuntimeError("NVRTC failed for %s:\n%s" % (arch, log.decode(errors="replace")))
        size = _check(nvrtc.nvrtcGetCUBINSize(prog))
        cubin = b" " * size
        _check(nvrtc.nvrtcGetCUBIN(prog, cubin))
        self.cubin = cubin
        self.module = _check(driver.cuModuleLoadData(cubin))

    def get_function(self, name):
        return _Function(self.module, name)


def _memoize(for_each_device=False):
    def deco(fn):
        return functools.lru_cache(maxsize=None)(fn)
    return deco


def install():
    """Register the stand-in as `cupy` (no-op when a real cupy is importable)."""
    if "cupy" in sys.modules and not getattr(sys.modules["cupy"], "_mi_stub", False):
        return sys.modules["cupy"]
    cupy = types.ModuleType("cupy")
    cupy._mi_stub = True
    cupy.util = types.SimpleNamespace(memoize=_memoize)
    cupy.cuda = types.SimpleNamespace(compile_with_cache=lambda src, *a, **k: _Module(src))
    sys.modules["cupy"] = cupy
    return cupy
