"""The CUDA path against the reference's OWN GPU code, live on the box: its separable-convolution kernel strings
(reference sepconv/sepconv_op/sepconv.py:5-30, 138-190) compiled with NVRTC for sm_100a and launched through the
`cupy` stand-in of baseline/cupy_nvrtc.py, and its whole `run_train_iter` (cuDNN fp32 + those kernels) in a
subprocess.  Needs the staged, git-ignored copy of the reference under baseline/_ref/ (it travels with the snapshot);
skipped when that copy is absent."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

from helpers import make_args

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "sepconv", "sepconv_op", "sepconv.py")),
                               reason="baseline/_ref (staged copy of the reference) is not present")


@pytest.fixture(scope="module")
def ref_sepconv():
    sys.path.insert(0, ROOT)
    from baseline import cupy_nvrtc
    cupy_nvrtc.install()
    spec = importlib.util.spec_from_file_location("ref_sepconv_op", os.path.join(REF, "sepconv", "sepconv_op",
                                                                              "sepconv.py"))
    mod = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod.FunctionSepconv


def _both(fn_ref, fn_own, inp, v, h, go):
    outs = []
    for fn in (fn_ref, fn_own):
        vv, hh = v.clone().requires_grad_(True), h.clone().requires_grad_(True)
        out = fn.apply(inp, vv, hh)
        gv, gh = torch.autograd.grad(out, (vv, hh), go)
        outs.append((out.detach(), gv, gh))
    torch.cuda.synchronize()
    return outs


@needs_ref
@pytest.mark.parametrize("n,ho,wo", [(2, 256, 448), (1, 37, 53)])
def test_sepconv_op_against_reference_nvrtc_kernels(cuda_ops, ref_sepconv, n, ho, wo):
    """FunctionSepconv.apply (drop-in op, csrc/sepconv.cu) vs the reference's kernels at F=51 on the BASELINE window
    (and a ragged one): output, gradVertical, gradHorizontal to 2e-5 of the scale."""
    from meta_interpolation_b200.sepconv.sepconv_op.sepconv import FunctionSepconv
    g = torch.Generator(device="cuda").manual_seed(0)
    inp = torch.rand(n, 3, ho + 50, wo + 50, device="cuda", generator=g)
    v = torch.randn(n, 51, ho, wo, device="cuda", generator=g) * 0.2
    h = torch.randn(n, 51, ho, wo, device="cuda", generator=g) * 0.2
    go = torch.randn(n, 3, ho, wo, device="cuda", generator=g)
    (ro, rgv, rgh), (oo, ogv, ogh) = _both(ref_sepconv, FunctionSepconv, inp, v, h, go)
    for name, a, b in (("out", oo, ro), ("gradVertical", ogv, rgv), ("gradHorizontal", ogh, rgh)):
        scale = b.abs().max().item()
        assert scale > 0
        assert (a - b).abs().max().item() <= 2e-5 * scale, name


@needs_ref
def test_fused_canvas_geometry_against_reference_padding_chain(cuda_ops, ref_sepconv):
    """The form the backbone uses (sepconv/model.py:254-266, 346-350): modulePaddingInput, modulePad and the output
    crop folded into the kernel (gy0=gx0=25, iy0=ix0=-25 on the raw frame) vs the reference's op on the explicitly
    padded 434x562 frame followed by the crop; gradients are compared inside the window (zero outside it)."""
    ops = cuda_ops
    n, H, W = 2, 256, 448
    g = torch.Generator(device="cuda").manual_seed(1)
    frame = torch.rand(n, 3, H, W, device="cuda", generator=g)
    v = torch.randn(n, 51, 384, 512, device="cuda", generator=g) * 0.2
    h = torch.randn(n, 51, 384, 512, device="cuda", generator=g) * 0.2
    go = torch.randn(n, 3, H, W, device="cuda", generator=g)
    pad_in = torch.nn.ReplicationPad2d([25, 39, 25, 103])
    pad = torch.nn.ReplicationPad2d([25, 25, 25, 25])
    vv, hh = v.clone().requires_grad_(True), h.clone().requires_grad_(True)
    full = ref_sepconv.apply(pad(pad_in(frame)).contiguous(), vv, hh)
    ref_out = full[:, :, 25:25 + H, 25:25 + W]
    rgv, rgh = torch.autograd.grad(ref_out, (vv, hh), go)
    vn = ops.empty_act(n, 384, 512, 51); vn.copy_(v.permute(0, 2, 3, 1))
    hn = ops.empty_act(n, 384, 512, 51); hn.copy_(h.permute(0, 2, 3, 1))
    ws = ops.sepconv_planar(n, H, W, 51)          # the four-pixels-per-thread kernels (what the backbone runs)
    out = ops.sepconv_fwd(frame, vn, hn, H, W, 25, 25, -25, -25, planar=ws)
    gv, gh = ops.zeros_act(n, 384, 512, 51), ops.zeros_act(n, 384, 512, 51)
    ops.sepconv_bwd(frame, vn, hn, go, gv, gh, 25, 25, -25, -25, planar=ws, planar_valid=True,
                    planar_grad=ops.sepconv_planar(n, H, W, 51))
    torch.cuda.synchronize()
    assert (out - ref_out).abs().max().item() <= 2e-5 * ref_out.abs().max().item()
    for name, a, b in (("gV", gv, rgv), ("gH", gh, rgh)):
        assert (a.permute(0, 3, 1, 2) - b).abs().max().item() <= 2e-5 * b.abs().max().item(), name


@needs_ref
def test_full_size_train_iter_against_live_reference_gpu(cuda_ops, tmp_path):
    """BASELINE configs[1] (sepconv, 256x448, K=5, LSLR-SGD), two tasks: the reference's own GPU path (cuDNN with
    TF32 off + its cupy kernels, subprocess on one visible GPU) and the graph-captured fast path on identical
    frames and seeded init: loss, predictions, PSNR (north_star: |dPSNR| < 0.01 dB)."""
    from bench import synthetic_septuplets
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    dump = str(tmp_path / "ref.pt")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    res = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "reference_gpu.py"), "--model", "sepconv",
                          "--batch", "2", "--steps", "1", "--warmup", "0", "--no-tf32", "--dump", dump],
                         capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    ref = torch.load(dump, weights_only=False)
    system = SceneAdaptiveInterpolation(make_args(cuda=True, batch_size=2, number_of_training_steps_per_iter=5),
                                        ops=cuda_ops)
    frames = [f.cuda() for f in synthetic_septuplets(2, 100)]
    losses, preds, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
    torch.cuda.synchronize()
    dl = abs(float(losses["loss"]) - ref["loss"])
    dp = (torch.cat(preds).cpu() - ref["preds"]).abs().max().item()
    dpsnr = abs(metrics["psnr"].avg - ref["psnr"])
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_live_reference_gpu.json"), "w") as f:
        json.dump(dict(reference=line, d_loss=dl, d_pred_maxabs=dp, d_psnr=dpsnr, psnr=ref["psnr"]), f)
    assert dl <= 5e-4 and dp <= 5e-3 and dpsnr < 0.01, (dl, dp, dpsnr)
