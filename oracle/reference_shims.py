"""Run the UNMODIFIED reference on CPU in the build container (test infrastructure).

Only ``oracle/make_golden.py`` and the container-only pin tests use this; the
GPU box has no ``/root/reference``.  No reference file is modified or copied:
the reference is imported from where it lies and monkey-patched in memory
(SURVEY.md Appendix C):

1. fake ``cupy`` module so ``sepconv/sepconv_op/sepconv.py:1`` imports, and
   ``FunctionSepconv.apply`` -> ``oracle.sepconv_op.FunctionSepconvCPU.apply``
   (the reference has no CPU path, sepconv.py:293-294);
2. ``ReduceLROnPlateau(verbose=...)`` accepted (meta_learning_system.py:144);
3. ``utils.load_checkpoint`` no-op and ``args.resume=True`` so no pretrained
   file is read (meta_learning_system.py:52,154-156);
4. ``Tensor.cuda`` / ``Module.cuda`` identity on CPU (sepconv/model.py:263).
"""
import os
import sys
import types
import contextlib
import io

import torch

REFERENCE_ROOT = os.environ.get("MI_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "meta_learning_system.py"))


_installed = False


def install():
    """Idempotently install the shims and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    # (1) fake cupy
    if "cupy" not in sys.modules:
        cupy = types.ModuleType("cupy")
        cupy.util = types.SimpleNamespace(memoize=lambda for_each_device=False: (lambda fn: fn))
        cupy.cuda = types.SimpleNamespace()
        sys.modules["cupy"] = cupy
    # (2) scheduler kwarg
    import torch.optim.lr_scheduler as lrs
    if not getattr(lrs.ReduceLROnPlateau, "_mi_shim", False):
        _orig = lrs.ReduceLROnPlateau

        class ReduceLROnPlateau(_orig):
            _mi_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        lrs.ReduceLROnPlateau = ReduceLROnPlateau
        torch.optim.lr_scheduler.ReduceLROnPlateau = ReduceLROnPlateau
    # (4) .cuda() identity on CPU
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)  # sepconv/model.py falls back to a cwd-relative sys.path entry
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import utils as ref_utils
            ref_utils._original_load_checkpoint = ref_utils.load_checkpoint   # kept for the interchange test
            ref_utils.load_checkpoint = lambda *a, **k: None  # (3)
            import sepconv.model  # noqa: F401
            from sepconv.sepconv_op import sepconv as ref_sepconv
            from oracle.sepconv_op import FunctionSepconvCPU
            ref_sepconv.FunctionSepconv.apply = FunctionSepconvCPU.apply  # (1)
    finally:
        os.chdir(cwd)
    _installed = True


def make_args(**over):
    """argparse.Namespace exactly as the reference's config.get_args() builds it
    (config.py:79-89) with an empty command line, then overridden."""
    install()
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        import config as ref_config
        args, _ = ref_config.get_args()
    finally:
        sys.argv = argv
    args = type(args)(**vars(args))
    args.cuda = False
    args.num_gpu = 0
    args.resume = True
    args.batch_size = 1
    for k, v in over.items():
        setattr(args, k, v)
    return args


def build_system(**over):
    """Construct the reference's SceneAdaptiveInterpolation on CPU."""
    install()
    args = make_args(**over)
    with contextlib.redirect_stdout(io.StringIO()):
        import meta_learning_system as ref_mls
        system = ref_mls.SceneAdaptiveInterpolation(args)
    return system, args
