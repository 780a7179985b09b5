"""tcgen05 (TF32) convolution engine against the fp32 CPU reference and the on-device SIMT engine.

Tolerance: TF32 operands carry a 10-bit mantissa (2^-11 relative rounding per operand); with fp32
accumulation the result differs from fp32 by ~1e-3 of the output scale.  Stated bar: 4e-3 * max|ref|."""
import pytest
import torch

from oracle.ops_ref import RefOps
from meta_interpolation_b200.ops import ACT_NONE, ACT_RELU, ENGINE_SIMT, ENGINE_TC, WG_STORE, WgradSpec, pad4
from test_kernels_gpu import act_pair, weight_pair, close

pytestmark = pytest.mark.gpu
REF = RefOps()
TF32_TOL = 4e-3

TC_SHAPES = [
    # n, h, w, cin, cout, k
    (2, 16, 24, 32, 32, 3), (1, 12, 16, 64, 51, 3), (1, 16, 16, 51, 51, 3), (2, 6, 8, 128, 64, 3),
    (1, 24, 32, 256, 512, 3), (1, 9, 14, 64, 64, 3), (1, 10, 12, 20, 32, 7), (2, 8, 16, 64, 64, 5),
    (1, 48, 64, 64, 64, 3), (1, 8, 8, 192, 16, 1),
    # persistent halo kernel: ragged tiles, and more tiles than SMs (TMEM double buffering, stage wrap-around)
    (1, 17, 13, 40, 64, 3), (2, 200, 208, 64, 64, 3), (1, 130, 300, 32, 32, 3),
    # stem: 6 input channels (32-channel TMA box over a 6-channel tensor, zero-filled tail)
    (1, 8, 8, 6, 32, 3), (2, 40, 48, 6, 32, 3),
    # filter-column weight-gradient kernel: Cout = 128 (four dY boxes), ragged Cout / Cin, one-tile layer,
    # many tiles per CTA (stage wrap-around), width not a multiple of the 8-pixel tile
    (2, 64, 72, 32, 128, 3), (1, 30, 50, 64, 100, 3), (1, 8, 8, 64, 64, 3), (2, 137, 233, 64, 64, 3),
    (2, 96, 128, 51, 51, 3),
    # streamed-weights halo kernel (> 64 channels): two tiles per item, one tile per item with 128- and 64-wide cout
    # tiles, odd tile counts (the last group re-reads a tile), ragged channels, the deepest SepConv layers
    (2, 96, 128, 128, 128, 3), (2, 48, 64, 256, 256, 3), (2, 24, 32, 512, 512, 3), (2, 12, 16, 512, 512, 3),
    (1, 40, 72, 128, 64, 3), (3, 33, 41, 96, 160, 3), (2, 96, 136, 64, 128, 3), (1, 50, 70, 200, 300, 3),
    # 5x5 / 7x7 layers (superslomo, voxelflow): filter-column weight gradient with 5 / 7 rows stacked along N
    (2, 40, 56, 32, 32, 7), (1, 33, 47, 64, 64, 5), (1, 32, 40, 6, 32, 7), (1, 24, 24, 128, 64, 5), (2, 16, 24, 20, 32, 7),
]


def _require_tc(ops):
    if not ops.lib.mi_tc_available():
        pytest.skip("tcgen05 path not available on this device")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_fprop_tc(cuda_ops, shape):
    _require_tc(cuda_ops)
    n, h, w, cin, cout, k = shape
    xc, xd = act_pair(cuda_ops, n, h, w, cin, 41)
    wc, wd = weight_pair(cuda_ops, cout, cin, k, 42)
    bc = torch.rand(cout) - 0.5
    yc = REF.conv_fprop(xc, wc, bc, ACT_RELU)
    yd = cuda_ops.conv_fprop(xd, wd, bc.cuda(), ACT_RELU, engine=ENGINE_TC)
    close(yd, yc, TF32_TOL, "tc fprop")
    ys = cuda_ops.conv_fprop(xd, wd, bc.cuda(), ACT_RELU, engine=ENGINE_SIMT)
    close(yd, ys.cpu(), TF32_TOL, "tc vs simt")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_dgrad_tc(cuda_ops, shape):
    _require_tc(cuda_ops)
    n, h, w, cin, cout, k = shape
    if cin < 16:
        pytest.skip("dgrad output channels below the tensor-core threshold")
    dyc, dyd = act_pair(cuda_ops, n, h, w, cout, 43)
    wc, wd = weight_pair(cuda_ops, cout, cin, k, 44)
    mc, md = act_pair(cuda_ops, n, h, w, cin, 45)
    dxc = REF.conv_dgrad(dyc, wc, mask_y=mc, mask_act=ACT_RELU)
    dxd = cuda_ops.conv_dgrad(dyd, wd, mask_y=md, mask_act=ACT_RELU, engine=ENGINE_TC)
    close(dxd, dxc, TF32_TOL, "tc dgrad")
    acc_c, acc_d = act_pair(cuda_ops, n, h, w, cin, 46)
    REF.conv_dgrad(dyc, wc, out=acc_c, accumulate=True)
    cuda_ops.conv_dgrad(dyd, wd, out=acc_d, accumulate=True, engine=ENGINE_TC)
    close(acc_d, acc_c, TF32_TOL, "tc dgrad accumulate")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_wgrad_tc(cuda_ops, shape):
    _require_tc(cuda_ops)
    n, h, w, cin, cout, k = shape
    xc, xd = act_pair(cuda_ops, n, h, w, cin, 47)
    dyc, dyd = act_pair(cuda_ops, n, h, w, cout, 48)
    ld = pad4(cin)
    gwc, gwd = REF.empty_weight(cout, cin, k), cuda_ops.empty_weight(cout, cin, k)
    gbc, gbd = torch.zeros(cout), torch.zeros(cout, device="cuda")
    REF.conv_wgrad(xc, dyc, k, ld, WgradSpec(WG_STORE, grad_w=gwc, grad_b=gbc))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_STORE, grad_w=gwd, grad_b=gbd), engine=ENGINE_TC)
    close(gwd, gwc, TF32_TOL, "tc wgrad w")
    close(gbd, gbc, 1e-4, "tc wgrad b")


def test_large_canvas_tc_vs_simt(cuda_ops):
    """BASELINE canvas (384x512): the two engines agree on the 51->51 full-resolution layer."""
    _require_tc(cuda_ops)
    ops = cuda_ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = ops.empty_act(1, 384, 512, 51)
    x.copy_(torch.rand(1, 384, 512, 51, device="cuda", generator=g) - 0.5)
    w = ops.empty_weight(51, 51, 3)
    w.copy_((torch.rand(51, 3, 3, 51, device="cuda", generator=g) - 0.5) * 0.1)
    b = torch.rand(51, device="cuda", generator=g)
    yt = ops.conv_fprop(x, w, b, engine=ENGINE_TC)
    ys = ops.conv_fprop(x, w, b, engine=ENGINE_SIMT)
    d = (yt - ys).abs().max().item()
    assert d <= TF32_TOL * ys.abs().max().item(), d


POISON_SHAPES = [
    # n, h, w, cin, cout  (3x3): U-Net pyramids at miniature sizes -- few tiles, many channel blocks, concat inputs
    (1, 64, 128, 6, 32), (1, 64, 128, 32, 32), (1, 32, 64, 32, 64), (1, 16, 32, 64, 128), (1, 8, 16, 128, 256),
    (1, 4, 8, 256, 512), (1, 4, 8, 512, 512), (1, 8, 16, 1024, 512), (1, 16, 32, 512, 256), (1, 32, 64, 256, 128),
    (1, 64, 128, 128, 64), (1, 64, 128, 64, 32), (2, 8, 8, 96, 160), (1, 16, 16, 200, 300), (1, 24, 40, 51, 51),
    (1, 64, 128, 16, 32), (1, 64, 128, 10, 32), (1, 20, 24, 48, 48),
    # the RRIN pyramid on a 128x128 canvas (batch 1), including its 9- and 10-channel inputs (row stride 12)
    (1, 8, 8, 256, 512), (1, 8, 8, 512, 512), (1, 16, 16, 128, 256), (1, 16, 16, 512, 256), (1, 32, 32, 256, 128),
    (1, 64, 64, 128, 64), (1, 128, 128, 9, 32), (1, 128, 128, 10, 32), (1, 128, 128, 64, 32), (1, 128, 128, 32, 32),
]


@pytest.mark.parametrize("shape", POISON_SHAPES)
def test_wgrad_never_reads_stale_workspace(cuda_ops, shape):
    """Every split-K partial the finishing kernel reads must have been written by the launch itself: the workspace is
    poisoned with NaN before the call (a hole shows up as NaN / garbage only when the allocator hands back dirty
    memory, i.e. depending on what ran before)."""
    _require_tc(cuda_ops)
    n, h, w, cin, cout = shape
    k = 3
    xc, xd = act_pair(cuda_ops, n, h, w, cin, 61)
    dyc, dyd = act_pair(cuda_ops, n, h, w, cout, 62)
    ld = pad4(cin)
    for t, c in ((xd, cin), (dyd, cout)):        # pad lanes of the 4-padded rows are poisoned too
        if pad4(c) != c:
            t.as_strided((n, h, w, pad4(c)), (h * w * pad4(c), w * pad4(c), pad4(c), 1),
                         t.storage_offset())[..., c:].fill_(float("nan"))
    need = cuda_ops.lib.mi_conv2d_wgrad_workspace(n, h, w, cin, cout, k, 0)
    ws = cuda_ops.workspace(need)
    ws.view(torch.float32)[: ws.numel() // 4].fill_(float("nan"))
    gwc, gwd = REF.empty_weight(cout, cin, k), cuda_ops.empty_weight(cout, cin, k)
    gbc, gbd = torch.zeros(cout), torch.zeros(cout, device="cuda")
    REF.conv_wgrad(xc, dyc, k, ld, WgradSpec(WG_STORE, grad_w=gwc, grad_b=gbc))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_STORE, grad_w=gwd, grad_b=gbd))
    assert torch.isfinite(gwd).all() and torch.isfinite(gbd).all()
    close(gwd, gwc, TF32_TOL, "wgrad w")
    close(gbd, gbc, 1e-4, "wgrad b")
