"""Every CUDA kernel of libmi_b200 against its CPU reference (oracle/ops_ref.py) on the same inputs."""
import pytest
import torch

from oracle.ops_ref import RefOps
from meta_interpolation_b200.ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ENGINE_SIMT, ENGINE_TC,
                                         WG_ACCUM, WG_SGD_SCALAR, WG_SGD_TENSOR, WG_STORE, WgradSpec, pad4)

pytestmark = pytest.mark.gpu
REF = RefOps()


def act_pair(ops, n, h, w, c, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    cpu = REF.empty_act(n, h, w, c)
    cpu.copy_((torch.rand(n, h, w, c, generator=g) - 0.5) * 2 * scale)
    dev = ops.empty_act(n, h, w, c)
    dev.copy_(cpu.cuda())
    return cpu, dev


def weight_pair(ops, cout, cin, k, seed):
    g = torch.Generator().manual_seed(seed)
    cpu = REF.empty_weight(cout, cin, k)
    cpu.copy_((torch.rand(cout, k, k, cin, generator=g) - 0.5) * (2.0 / (cin * k * k) ** 0.5))
    dev = ops.empty_weight(cout, cin, k)
    dev.copy_(cpu.cuda())
    return cpu, dev


def close(dev, cpu, tol, what=""):
    d = (dev.detach().cpu() - cpu).abs().max().item()
    s = cpu.abs().max().item()
    assert d <= tol * max(s, 1e-6), "%s: max|diff| %.3e vs scale %.3e" % (what, d, s)


CONV_SHAPES = [
    # n, h, w, cin, cout, k
    (1, 8, 8, 6, 32, 3), (2, 16, 24, 32, 32, 3), (1, 12, 16, 64, 51, 3), (1, 16, 16, 51, 51, 3),
    (2, 6, 8, 128, 64, 3), (1, 9, 7, 5, 3, 5), (1, 10, 12, 20, 32, 7), (1, 4, 4, 192, 12, 1), (1, 3, 5, 7, 130, 3),
    # few-output-channel heads (flow / mask / RGB): the per-pixel forward and per-(tap,cin) weight-gradient kernels
    (2, 16, 24, 32, 5, 3), (1, 20, 28, 32, 4, 3), (1, 12, 12, 64, 3, 5), (2, 9, 11, 32, 2, 3), (1, 40, 33, 33, 8, 3),
    (1, 70, 90, 32, 4, 3), (2, 12, 20, 32, 10, 3), (1, 16, 16, 32, 9, 3), (1, 10, 14, 64, 3, 5), (1, 6, 6, 250, 2, 3),
    # feature maps of a few pixels (CAIN's channel attention: 1x1 convs on [N, C, 1, 1]): one warp per output element
    (2, 1, 1, 192, 12, 1), (2, 1, 1, 12, 192, 1), (1, 1, 1, 64, 4, 1), (4, 2, 2, 70, 33, 3), (1, 3, 5, 40, 7, 3),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("act", [ACT_NONE, ACT_RELU, ACT_LEAKY])
def test_conv_fprop_simt(cuda_ops, shape, act):
    n, h, w, cin, cout, k = shape
    xc, xd = act_pair(cuda_ops, n, h, w, cin, 1)
    wc, wd = weight_pair(cuda_ops, cout, cin, k, 2)
    bc = torch.rand(cout) - 0.5
    yc = REF.conv_fprop(xc, wc, bc, act, 0.1)
    yd = cuda_ops.conv_fprop(xd, wd, bc.cuda(), act, 0.1, engine=ENGINE_SIMT)
    close(yd, yc, 2e-5, "fprop")


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_dgrad_simt(cuda_ops, shape):
    n, h, w, cin, cout, k = shape
    dyc, dyd = act_pair(cuda_ops, n, h, w, cout, 3)
    wc, wd = weight_pair(cuda_ops, cout, cin, k, 4)
    mc, md = act_pair(cuda_ops, n, h, w, cin, 5)
    dxc = REF.conv_dgrad(dyc, wc, mask_y=mc, mask_act=ACT_RELU)
    dxd = cuda_ops.conv_dgrad(dyd, wd, mask_y=md, mask_act=ACT_RELU, engine=ENGINE_SIMT)
    close(dxd, dxc, 2e-5, "dgrad")
    # accumulate
    acc_c, acc_d = act_pair(cuda_ops, n, h, w, cin, 6)
    REF.conv_dgrad(dyc, wc, out=acc_c, accumulate=True)
    cuda_ops.conv_dgrad(dyd, wd, out=acc_d, accumulate=True, engine=ENGINE_SIMT)
    close(acc_d, acc_c, 2e-5, "dgrad accumulate")


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_wgrad_simt_modes(cuda_ops, shape):
    n, h, w, cin, cout, k = shape
    xc, xd = act_pair(cuda_ops, n, h, w, cin, 7)
    dyc, dyd = act_pair(cuda_ops, n, h, w, cout, 8)
    wc, wd = weight_pair(cuda_ops, cout, cin, k, 9)
    bc = torch.rand(cout)
    bd = bc.cuda()
    ld = pad4(cin)
    # STORE
    gwc, gwd = REF.empty_weight(cout, cin, k), cuda_ops.empty_weight(cout, cin, k)
    gbc, gbd = torch.zeros(cout), torch.zeros(cout, device="cuda")
    REF.conv_wgrad(xc, dyc, k, ld, WgradSpec(WG_STORE, grad_w=gwc, grad_b=gbc))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_STORE, grad_w=gwd, grad_b=gbd), engine=ENGINE_SIMT)
    close(gwd, gwc, 3e-5, "wgrad store w")
    close(gbd, gbc, 3e-5, "wgrad store b")
    # ACCUM with scale on top of the stored gradient
    REF.conv_wgrad(xc, dyc, k, ld, WgradSpec(WG_ACCUM, scale=0.25, grad_w=gwc, grad_b=gbc))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_ACCUM, scale=0.25, grad_w=gwd, grad_b=gbd), engine=ENGINE_SIMT)
    close(gwd, gwc, 3e-5, "wgrad accum w")
    # fused LSLR SGD update, in place, with gsum
    lr = torch.tensor([0.05])
    sc = WgradSpec(WG_SGD_SCALAR, w_in=wc, b_in=bc, w_out=wc, b_out=bc, lr_w=lr, lr_b=lr,
                   gsum_w=REF.empty_weight(cout, cin, k), gsum_b=torch.zeros(cout))
    sd = WgradSpec(WG_SGD_SCALAR, w_in=wd, b_in=bd, w_out=wd, b_out=bd, lr_w=lr.cuda(), lr_b=lr.cuda(),
                   gsum_w=cuda_ops.empty_weight(cout, cin, k), gsum_b=torch.zeros(cout, device="cuda"))
    REF.conv_wgrad(xc, dyc, k, ld, sc)
    cuda_ops.conv_wgrad(xd, dyd, k, ld, sd, engine=ENGINE_SIMT)
    close(wd, wc, 3e-5, "fused sgd w")
    close(bd, bc, 3e-5, "fused sgd b")
    close(sd.gsum_w, sc.gsum_w, 3e-5, "gsum")
    # Meta-SGD (per-element alpha)
    ac, ad = weight_pair(cuda_ops, cout, cin, k, 10)
    ab = torch.rand(cout)
    w2c, w2d = REF.empty_weight(cout, cin, k), cuda_ops.empty_weight(cout, cin, k)
    b2c, b2d = torch.zeros(cout), torch.zeros(cout, device="cuda")
    REF.conv_wgrad(xc, dyc, k, ld, WgradSpec(WG_SGD_TENSOR, w_in=wc, b_in=bc, w_out=w2c, b_out=b2c, lr_w=ac, lr_b=ab))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_SGD_TENSOR, w_in=wd, b_in=bd, w_out=w2d, b_out=b2d, lr_w=ad,
                                                  lr_b=ab.cuda()), engine=ENGINE_SIMT)
    close(w2d, w2c, 3e-5, "metasgd w")
    # the fused update can also emit the updated weight in the rotated layout dgrad reads
    w3d, wt3 = cuda_ops.empty_weight(cout, cin, k), cuda_ops.empty_weight(cin, cout, k)
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_SGD_SCALAR, w_in=wd, b_in=bd, w_out=w3d, b_out=b2d,
                                                  lr_w=lr.cuda(), lr_b=lr.cuda(), wt_out=wt3), engine=ENGINE_SIMT)
    close(wt3, cuda_ops.weight_to_dgrad(w3d, rnd=False).cpu(), 0.0,
          "fused rotated weight == weight_to_dgrad(updated weight)")
    # ... and, with a TF32-rounded fprop copy requested (wr_out), both copies are on the TF32 grid while the master
    # copy w_out stays exact (TF32 operand convention, include/mi_b200.h)
    w4d, wt4, wr4 = (cuda_ops.empty_weight(cout, cin, k), cuda_ops.empty_weight(cin, cout, k),
                     cuda_ops.empty_weight(cout, cin, k))
    cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_SGD_SCALAR, w_in=wd, b_in=bd, w_out=w4d, b_out=b2d,
                                                  lr_w=lr.cuda(), lr_b=lr.cuda(), wt_out=wt4, wr_out=wr4),
                        engine=ENGINE_SIMT)
    close(w4d, w3d.cpu(), 0.0, "master copy unchanged by wr_out")
    close(wr4, REF.round_tf32(w3d.cpu().contiguous()), 0.0, "wr_out == rn_tf32(w_out)")
    close(wt4, cuda_ops.weight_to_dgrad(w3d, rnd=True).cpu(), 0.0, "rotated copy rounded with wr_out")
    with pytest.raises(Exception):      # only the SGD modes own an updated weight to rotate
        cuda_ops.conv_wgrad(xd, dyd, k, ld, WgradSpec(WG_STORE, grad_w=gwd, grad_b=gbd, wt_out=wt3), engine=ENGINE_SIMT)
    # pad lanes of the KRSC storage stay zero
    if ld != cin:
        full = w2d.as_strided((cout, k, k, ld), (k * k * ld, k * ld, ld, 1), w2d.storage_offset())
        assert float(full[..., cin:].abs().max()) == 0.0


def test_weight_to_dgrad(cuda_ops):
    wc, wd = weight_pair(cuda_ops, 7, 5, 3, 11)
    close(cuda_ops.weight_to_dgrad(wd, rnd=False), REF.weight_to_dgrad(wc, rnd=False), 0.0, "weight_to_dgrad")
    close(cuda_ops.weight_to_dgrad(wd, rnd=True), REF.weight_to_dgrad(wc, rnd=True), 0.0, "weight_to_dgrad, rounded")


@pytest.mark.parametrize("c", [3, 32, 51])
def test_pool_upsample_add_copy_act(cuda_ops, c):
    n, h, w = 2, 8, 12
    xc, xd = act_pair(cuda_ops, n, h, w, c, 12)
    close(cuda_ops.avgpool_fwd(xd), REF.avgpool_fwd(xc), 1e-6, "avgpool")
    close(cuda_ops.maxpool_fwd(xd), REF.maxpool_fwd(xc), 0.0, "maxpool")
    gyc, gyd = act_pair(cuda_ops, n, h // 2, w // 2, c, 13)
    for acc in (False, True):
        dc, dd = act_pair(cuda_ops, n, h, w, c, 14)
        REF.avgpool_bwd(gyc, dc, acc)
        cuda_ops.avgpool_bwd(gyd, dd, acc)
        close(dd, dc, 1e-6, "avgpool bwd")
        dc, dd = act_pair(cuda_ops, n, h, w, c, 15)
        REF.maxpool_bwd(xc, gyc, dc, acc)
        cuda_ops.maxpool_bwd(xd, gyd, dd, acc)
        close(dd, dc, 1e-6, "maxpool bwd")
    for align in (True, False):
        close(cuda_ops.upsample_fwd(xd, align), REF.upsample_fwd(xc, align), 2e-6, "upsample")
        guc, gud = act_pair(cuda_ops, n, 2 * h, 2 * w, c, 16)
        for acc in (False, True):
            dc, dd = act_pair(cuda_ops, n, h, w, c, 17)
            REF.upsample_bwd(guc, dc, align, acc)
            cuda_ops.upsample_bwd(gud, dd, align, acc)
            close(dd, dc, 2e-6, "upsample bwd align=%s" % align)
    # activation derivative of the upsampled tensor fused into the upsample backward
    for act in (ACT_RELU, ACT_LEAKY):
        for align in (True, False):
            guc, gud = act_pair(cuda_ops, n, 2 * h, 2 * w, c, 24)
            dc, dd = act_pair(cuda_ops, n, h, w, c, 25)
            REF.upsample_bwd(guc, dc, align, False, xc, act, 0.1)
            cuda_ops.upsample_bwd(gud, dd, align, False, xd, act, 0.1)
            close(dd, dc, 2e-6, "upsample bwd with fused act mask")
    yc, yd = act_pair(cuda_ops, n, h, w, c, 18)
    close(cuda_ops.add(xd, yd), REF.add(xc, yc), 1e-7, "add")
    for act in (ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH):
        gc, gd = act_pair(cuda_ops, n, h, w, c, 19)
        REF.act_bwd(gc, xc, act, 0.1)
        cuda_ops.act_bwd(gd, xd, act, 0.1)
        close(gd, gc, 1e-6, "act_bwd")
    # channel-slice copy into a wider concat buffer
    bufc, bufd = act_pair(cuda_ops, n, h, w, c + 8, 20)
    REF.copy(xc, bufc[..., 4:4 + c])
    cuda_ops.copy(xd, bufd[..., 4:4 + c])
    close(bufd, bufc, 0.0, "slice copy")


@pytest.mark.parametrize("mode,hw", [(0, (20, 28)), (1, (20, 28)), (0, (5, 3))])
def test_frames_to_canvas_and_windows(cuda_ops, mode, hw):
    h, w = hw
    g = torch.Generator().manual_seed(21)
    f0, f1 = torch.rand(2, 3, h, w, generator=g), torch.rand(2, 3, h, w, generator=g)
    pt, pl = (3, 2) if mode == 1 else (25, 25)
    ch, cw = (h + 8, w + 6) if mode == 1 else (128, 128)
    cc = REF.frames_to_canvas(f0, f1, ch, cw, pt, pl, mode)
    cd = cuda_ops.frames_to_canvas(f0.cuda(), f1.cuda(), ch, cw, pt, pl, mode)
    close(cd, cc, 0.0, "canvas")
    close(cuda_ops.nhwc_window_to_nchw(cd, pt, pl, h, w), REF.nhwc_window_to_nchw(cc, pt, pl, h, w), 0.0, "window")


@pytest.mark.parametrize("planar", [False, True])
@pytest.mark.parametrize("geom", [dict(taps=51, c=3, h=40, w=70), dict(taps=51, c=3, h=9, w=33),
                                  dict(taps=5, c=2, h=6, w=7)])
def test_sepconv_fused_geometry(cuda_ops, geom, planar):
    """Both kernel generations: filters read as NHWC rows (two pixels per thread) and, with the tap-planar workspace,
    four pixels per thread; the backward with the planar filters left by the forward and re-made by itself; rounded
    gradients on request."""
    taps, c, h, w = geom["taps"], geom["c"], geom["h"], geom["w"]
    if planar and taps != 51:
        pytest.skip("the planar kernels exist for F = 51, c = 3")
    pad = taps // 2
    gh, gw = h + 2 * pad + 3, w + 2 * pad + 5
    g = torch.Generator().manual_seed(22)
    frame = torch.rand(2, c, h, w, generator=g)
    vc, vd = act_pair(cuda_ops, 2, gh, gw, taps, 23, 0.2)
    hc, hd = act_pair(cuda_ops, 2, gh, gw, taps, 24, 0.2)
    oc = REF.sepconv_fwd(frame, vc, hc, h, w, pad, pad, -pad, -pad)
    ws = cuda_ops.sepconv_planar(2, h, w, taps) if planar else None
    od = cuda_ops.sepconv_fwd(frame.cuda(), vd, hd, h, w, pad, pad, -pad, -pad, planar=ws)
    close(od, oc, 2e-5, "sepconv fwd")
    go = torch.rand(2, c, h, w, generator=g) - 0.5
    gvc, ghc = REF.zeros_act(2, gh, gw, taps), REF.zeros_act(2, gh, gw, taps)
    gvd, ghd = cuda_ops.zeros_act(2, gh, gw, taps), cuda_ops.zeros_act(2, gh, gw, taps)
    REF.sepconv_bwd(frame, vc, hc, go, gvc, ghc, pad, pad, -pad, -pad)
    scratch = cuda_ops.sepconv_planar(2, h, w, taps) if planar else None
    cuda_ops.sepconv_bwd(frame.cuda(), vd, hd, go.cuda(), gvd, ghd, pad, pad, -pad, -pad, planar=ws,
                         planar_valid=planar, planar_grad=scratch)
    close(gvd, gvc, 3e-5, "sepconv gV")
    close(ghd, ghc, 3e-5, "sepconv gH")
    # the call zero-fills outside the window itself (MI_SEPCONV_ZERO_OUTSIDE): NaN-filled buffers come back equal to
    # the gradients written into zero-filled ones, pad lanes of the rows aside
    gv3, gh3 = cuda_ops.empty_act(2, gh, gw, taps), cuda_ops.empty_act(2, gh, gw, taps)
    gv3.fill_(float("nan")); gh3.fill_(float("nan"))
    cuda_ops.sepconv_bwd(frame.cuda(), vd, hd, go.cuda(), gv3, gh3, pad, pad, -pad, -pad, planar=ws,
                         planar_valid=planar, planar_grad=scratch, zero_outside=True)
    assert torch.equal(gv3.cpu(), gvd.cpu()) and torch.equal(gh3.cpu(), ghd.cpu()), "zero_outside changes the result"
    assert not torch.isnan(gv3).any() and not torch.isnan(gh3).any()
    if planar:      # a backward that has to transpose the filters itself, with rounded outputs
        ws2 = cuda_ops.sepconv_planar(2, h, w, taps)
        gv2, gh2 = cuda_ops.zeros_act(2, gh, gw, taps), cuda_ops.zeros_act(2, gh, gw, taps)
        cuda_ops.sepconv_bwd(frame.cuda(), vd, hd, go.cuda(), gv2, gh2, pad, pad, -pad, -pad, rnd=True, planar=ws2,
                             planar_valid=False, planar_grad=scratch)
        close(gv2, REF.round_tf32(gvc.clone().contiguous()), 3e-4, "sepconv gV, rounded")
        assert torch.equal(gv2.contiguous(), cuda_ops.round_tf32(gv2.contiguous().clone())) or not cuda_ops.tf32_rn
        # outside the window nothing was written
        assert float(gv2[:, :pad].abs().max()) == 0.0 and float(gh2[:, :, :pad].abs().max()) == 0.0


def test_function_sepconv_dropin_matches_reference_op(cuda_ops):
    """The reference's own calling convention (pre-padded NCHW input, [N,F,H,W] filters)."""
    from meta_interpolation_b200.sepconv.sepconv_op.sepconv import FunctionSepconv
    from oracle.sepconv_op import FunctionSepconvCPU
    g = torch.Generator().manual_seed(25)
    inp = torch.rand(1, 3, 16 + 50, 20 + 50, generator=g)
    v = (torch.rand(1, 51, 16, 20, generator=g) * 0.1).requires_grad_(True)
    h = (torch.rand(1, 51, 16, 20, generator=g) * 0.1).requires_grad_(True)
    vd, hd = v.detach().cuda().requires_grad_(True), h.detach().cuda().requires_grad_(True)
    out_c = FunctionSepconvCPU.apply(inp, v, h)
    out_d = FunctionSepconv.apply(inp.cuda(), vd, hd)
    close(out_d, out_c.detach(), 2e-5, "FunctionSepconv fwd")
    go = torch.rand(1, 3, 16, 20, generator=g)
    out_c.backward(go)
    out_d.backward(go.cuda())
    close(vd.grad, v.grad, 3e-5, "FunctionSepconv gV")
    close(hd.grad, h.grad, 3e-5, "FunctionSepconv gH")
    with pytest.raises(NotImplementedError):
        FunctionSepconv.apply(inp, v, h)   # CPU tensors: same error as the reference (sepconv.py:293-294)


@pytest.mark.parametrize("c", [3, 4, 5, 1])      # 16-byte pixels take the vector kernels (c <= 4), c = 5 the scalar ones
@pytest.mark.parametrize("variant,sx,sy", [(0, 1.0, 1.0), (1, -0.5, -0.5), (1, 0.5, 0.5)])
def test_warp(cuda_ops, variant, sx, sy, c):
    n, h, w = 2, 12, 17
    ic, idv = act_pair(cuda_ops, n, h, w, c, 26)
    fc, fd = act_pair(cuda_ops, n, h, w, 2, 27, 3.0 if variant == 0 else 0.6)
    close(cuda_ops.warp_fwd(idv, fd, variant, sx, sy), REF.warp_fwd(ic, fc, variant, sx, sy), 1e-5, "warp fwd")
    goc, god = act_pair(cuda_ops, n, h, w, c, 28)
    gfc, gfd = REF.zeros_act(n, h, w, 2), cuda_ops.zeros_act(n, h, w, 2)
    REF.warp_bwd(ic, fc, goc, gfc, variant, sx, sy)
    cuda_ops.warp_bwd(idv, fd, god, gfd, variant, sx, sy)
    close(gfd, gfc, 2e-4, "warp flow grad")
    REF.warp_bwd(ic, fc, goc, gfc, variant, sx, sy, accumulate=True)
    cuda_ops.warp_bwd(idv, fd, god, gfd, variant, sx, sy, accumulate=True)
    close(gfd, gfc, 2e-4, "warp flow grad, accumulated")


def test_identity_flow_is_half_pixel_shift(cuda_ops):
    # KAT: zero flow samples at x-0.5 -> average of the pixel and its upper-left neighbours, zeros outside
    img = cuda_ops.zeros_act(1, 4, 4, 1)
    img.fill_(1.0)
    out = cuda_ops.warp_fwd(img, cuda_ops.zeros_act(1, 4, 4, 2), 0)
    exp = torch.ones(4, 4)
    exp[0, :] = 0.5
    exp[:, 0] = 0.5
    exp[0, 0] = 0.25
    assert torch.allclose(out[0, :, :, 0].cpu(), exp, atol=1e-6)


@pytest.mark.parametrize("kind", [0, 1])
def test_loss_and_psnr(cuda_ops, kind):
    g = torch.Generator().manual_seed(29)
    p, t = torch.rand(2, 3, 9, 11, generator=g), torch.rand(2, 3, 9, 11, generator=g)
    lc, ld = torch.zeros(1), torch.zeros(1, device="cuda")
    gc, gd = torch.zeros_like(p), torch.zeros_like(p).cuda()
    REF.loss_fwd_bwd(p, t, kind, 2.0, lc, gc)
    cuda_ops.loss_fwd_bwd(p.cuda(), t.cuda(), kind, 2.0, ld, gd)
    close(ld, lc, 1e-5, "loss")
    close(gd, gc, 1e-6, "loss grad")
    sc, sd = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64, device="cuda")
    REF.psnr_accumulate(p, t, sc)
    cuda_ops.psnr_accumulate(p.cuda(), t.cuda(), sd)
    assert abs(float(sd) - float(sc)) <= 1e-9 * float(sc)


@pytest.mark.parametrize("rule", [0, 1, 2, 3])
def test_inner_update_rules(cuda_ops, rule):
    n = 3000
    g = torch.Generator().manual_seed(30)
    w, gr = torch.rand(n, generator=g), torch.rand(n, generator=g) - 0.5
    seg = torch.tensor([0, 1, -1], dtype=torch.int32)
    skip = torch.tensor([0, 1], dtype=torch.uint8)
    lr = torch.rand(2, 4, generator=g) * 0.1
    mc, vc = torch.rand(n, generator=g) * 0.1, torch.rand(n, generator=g) * 0.1
    md, vd = mc.cuda(), vc.cuda()
    for use_skip in (None, skip):
        oc, od = torch.zeros(n), torch.zeros(n, device="cuda")
        REF.inner_update(w, gr, oc, mc, vc, lr, False, 4, 2, seg, use_skip, rule, 2)
        cuda_ops.inner_update(w.cuda(), gr.cuda(), od, md, vd, lr.cuda(), False, 4, 2, seg.cuda(),
                              None if use_skip is None else use_skip.cuda(), rule, 2)
        close(od, oc, 2e-6, "inner rule %d" % rule)
        close(md, mc, 2e-6, "exp_avg")
    alpha = torch.rand(n, generator=g) * 0.1
    oc, od = torch.zeros(n), torch.zeros(n, device="cuda")
    REF.inner_update(w, gr, oc, mc, vc, alpha, True, 0, 0, seg, None, rule, 1)
    cuda_ops.inner_update(w.cuda(), gr.cuda(), od, md, vd, alpha.cuda(), True, 0, 0, seg.cuda(), None, rule, 1)
    close(od, oc, 2e-6, "per-element lr")


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_outer_step_matches_torch_optim(cuda_ops, kind):
    n = 5000
    g = torch.Generator().manual_seed(31)
    p0 = torch.rand(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = [torch.optim.SGD([ref], lr=1e-2), torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.99)),
           torch.optim.Adamax([ref], lr=1e-2, betas=(0.9, 0.999))][kind]
    pd = p0.cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in (1, 2, 3):
        gr = torch.rand(n, generator=g) - 0.5
        ref.grad = gr.clone()
        opt.step()
        cuda_ops.outer_step(pd, gr.cuda(), m, v, kind, 1e-2, 0.9, 0.999 if kind == 2 else 0.99, 1e-8, 0.0, step)
    close(pd, ref.detach(), 2e-6, "outer step")


def test_axpby_addcmul_segment_dot_fill(cuda_ops):
    n = 2500
    g = torch.Generator().manual_seed(32)
    a, b, c = (torch.rand(n, generator=g) for _ in range(3))
    yd = b.cuda()
    cuda_ops.axpby(a.cuda(), 0.5, yd, -2.0)
    close(yd, 0.5 * a - 2.0 * b, 1e-6, "axpby")
    yd = c.cuda()
    cuda_ops.addcmul(yd, -0.3, a.cuda(), b.cuda())
    close(yd, c - 0.3 * a * b, 1e-6, "addcmul")
    seg = torch.tensor([0, 2, -1], dtype=torch.int32)
    out = torch.zeros(3, device="cuda")
    cuda_ops.segment_dot(a.cuda(), b.cuda(), seg.cuda(), out)
    exp = torch.zeros(3)
    REF.segment_dot(a, b, seg, exp)
    close(out, exp, 1e-5, "segment_dot")
    cuda_ops.fill(yd, 3.0)
    assert float(yd.min()) == 3.0 == float(yd.max())


def test_segment_scale_and_segment_sum(cuda_ops):
    """L2F arena ops: theta' = gamma (.) theta, dL/dtheta += a * gamma' (.) G, per-tensor sums (embedding)."""
    n = 5 * 1024
    g = torch.Generator().manual_seed(40)
    x, y0 = torch.rand(n, generator=g) - 0.5, torch.rand(n, generator=g)
    seg = torch.tensor([0, 0, 1, -1, 2], dtype=torch.int32)
    gamma, mask = torch.rand(3, generator=g), torch.tensor([1.0, 0.0, 1.0])
    for m, acc, alpha in ((None, False, 1.0), (mask, True, 0.25)):
        yc, yd = y0.clone(), y0.clone().cuda()
        REF.segment_scale(x, gamma, seg, m, yc, alpha, acc)
        cuda_ops.segment_scale(x.cuda(), gamma.cuda(), seg.cuda(), None if m is None else m.cuda(), yd, alpha, acc)
        close(yd, yc, 1e-6, "segment_scale")
        assert torch.equal(yd[3 * 1024:4 * 1024].cpu(), y0[3 * 1024:4 * 1024])      # padding chunk untouched
    oc, od = torch.zeros(3), torch.zeros(3, device="cuda")
    REF.segment_dot(x, None, seg, oc)
    cuda_ops.segment_dot(x.cuda(), None, seg.cuda(), od)
    close(od, oc, 1e-5, "segment sum")


def test_bad_arguments_raise(cuda_ops):
    from meta_interpolation_b200._lib import MiB200Error
    x = cuda_ops.empty_act(1, 5, 5, 4)     # odd size cannot be pooled
    with pytest.raises(MiB200Error):
        cuda_ops.avgpool_fwd(x)
    w = cuda_ops.empty_weight(4, 4, 3)
    with pytest.raises(MiB200Error):
        cuda_ops.conv_fprop(cuda_ops.empty_act(1, 4, 4, 4), w, None, engine=ENGINE_TC)   # ineligible shape for tcgen05


@pytest.mark.parametrize("frame_hw", [(32, 32), (40, 56), (64, 64)])
@pytest.mark.parametrize("c", [51, 64])
def test_region_of_interest_upsample_and_crop(cuda_ops, frame_hw, c):
    """Windowed x2 upsample (weights of the full grid) and window copy used by the SepConv Subnet region of interest:
    equal to the full-canvas evaluation restricted to the window, and the backward is its exact adjoint."""
    from meta_interpolation_b200.sepconv.model import _axis_roi, canvas_size
    ch, cw = canvas_size(*frame_hw)
    hy0, hh, ly0, lh = _axis_roi(25, frame_hw[0], ch)
    hx0, hw, lx0, lw = _axis_roi(25, frame_hw[1], cw)
    n = 2
    fullc, fulld = act_pair(cuda_ops, n, ch // 2, cw // 2, c, 80)
    # crop
    cropc, cropd = REF.empty_act(n, lh, lw, c), cuda_ops.empty_act(n, lh, lw, c)
    REF.window_copy(fullc, (ly0, lx0), cropc, (0, 0), (lh, lw))
    cuda_ops.window_copy(fulld, (ly0, lx0), cropd, (0, 0), (lh, lw))
    close(cropd, cropc, 0.0, "window_copy")
    close(cropd, fullc[:, ly0:ly0 + lh, lx0:lx0 + lw, :], 0.0, "window_copy vs slice")
    # windowed upsample == window of the full upsample
    geo = dict(full_hw=(ch // 2, cw // 2), lo_origin=(ly0, lx0), hi_origin=(hy0, hx0))
    upd = cuda_ops.upsample_window_fwd(cropd, True, hi_hw=(hh, hw), **geo)
    full_up = cuda_ops.upsample_fwd(fulld, True)
    close(upd, full_up[:, hy0:hy0 + hh, hx0:hx0 + hw, :].cpu(), 0.0, "windowed upsample vs full upsample (bit exact)")
    close(upd, REF.upsample_window_fwd(cropc, True, hi_hw=(hh, hw), **geo), 2e-6, "windowed upsample vs reference")
    # adjoint
    gc, gd = act_pair(cuda_ops, n, hh, hw, c, 81)
    for acc in (False, True):
        dc, dd = act_pair(cuda_ops, n, lh, lw, c, 82)
        REF.upsample_window_bwd(gc, dc, True, acc, **geo)
        cuda_ops.upsample_window_bwd(gd, dd, True, acc, **geo)
        close(dd, dc, 3e-6, "windowed upsample bwd")
    # scatter back (adjoint of the crop) accumulates into the window only
    bigc, bigd = act_pair(cuda_ops, n, ch // 2, cw // 2, c, 83)
    REF.window_copy(cropc, (0, 0), bigc, (ly0, lx0), (lh, lw), accumulate=True)
    cuda_ops.window_copy(cropd, (0, 0), bigd, (ly0, lx0), (lh, lw), accumulate=True)
    close(bigd, bigc, 1e-6, "window accumulate")
    with pytest.raises(Exception):
        cuda_ops.window_copy(cropd, (0, 0), bigd, (ch // 2 - 1, 0), (lh, lw))


def test_subnet_region_of_interest_equals_full_canvas(cuda_ops):
    """SepConv forward + support-gradient with the Subnets on the region of interest vs on the full canvas: the
    prediction is bit-identical (same kernels, same values inside the window) and the gradients agree to rounding
    (split-K partitions differ with the buffer size)."""
    from meta_interpolation_b200.sepconv.model import MetaNetwork
    from oracle import backbones as bb
    bb.set_torch_seed(12345)
    net = MetaNetwork(ops=cuda_ops)
    g = torch.Generator().manual_seed(4)
    f0, f1, tgt = (torch.rand(1, 3, 72, 104, generator=g).cuda() for _ in range(3))
    outs, grads = [], []
    for roi in (True, False):
        net.SUBNET_ROI = roi
        fast = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
        out = net.forward(f0, f1, params=fast)
        gr = torch.autograd.grad((out - tgt).abs().mean(), list(fast.values()), allow_unused=True)
        outs.append(out.detach())
        grads.append(gr)
    net.SUBNET_ROI = True
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-6
    for a, b in zip(*grads):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a - b).abs().max().item() <= 2e-3 * max(b.abs().max().item(), 1e-8)


def _poisoned(ops, n, h, w, c, seed):
    """Activation whose pad lanes hold NaN (what a dirty allocator block looks like)."""
    cpu, dev = act_pair(ops, n, h, w, c, seed)
    if pad4(c) != c:
        dev.as_strided((n, h, w, pad4(c)), (h * w * pad4(c), w * pad4(c), pad4(c), 1),
                       dev.storage_offset())[..., c:].fill_(float("nan"))
    return cpu, dev


@pytest.mark.parametrize("c,width", [(2, 4), (6, 8), (3, 4), (18, 20), (51, 52)])
def test_ragged_slice_never_writes_neighbouring_channels(cuda_ops, c, width):
    """A [0:c] channel slice of a tensor whose width equals c rounded up to 4 is indistinguishable, by pointer and
    stride, from a c-channel tensor padded to 4: no kernel may store into the lanes past c (they are the next slice's
    channels), whatever the source's own pad lanes contain.  (Regression: a float4 tail store in the slice-gradient
    fan-in corrupted flow channels 2:4 of RRIN whenever the allocator returned dirty memory.)"""
    ops = cuda_ops
    n, h, w = 2, 6, 8
    srcc, srcd = _poisoned(ops, n, h, w, c, 90)
    othc, othd = _poisoned(ops, n, h, w, c, 91)

    def fresh():
        full = ops.zeros_act(n, h, w, width)
        full.fill_(7.0)
        return full

    def check(full, expect, what):
        assert torch.equal(full[..., c:].cpu(), torch.full((n, h, w, width - c), 7.0)), what + ": neighbours touched"
        close(full[..., :c], expect, 1e-6, what)

    full = fresh()
    ops.copy(srcd, full[..., :c], accumulate=True)
    check(full, 7.0 + srcc, "copy accumulate")
    full = fresh()
    ops.copy(srcd, full[..., :c], accumulate=False)
    check(full, srcc, "copy")
    full = fresh()
    ops.add(srcd, othd, out=full[..., :c])
    check(full, srcc + othc, "add")
    full = fresh()
    ops.affine(srcd, 0.5, 0.25, out=full[..., :c])
    check(full, 0.5 * srcc + 0.25, "affine")
    full = fresh()
    ops.act_fwd(srcd, ACT_TANH, 0.0, out=full[..., :c])
    check(full, torch.tanh(srcc), "act_fwd")
    big = ops.zeros_act(n, 2 * h, 2 * w, width)
    big.fill_(7.0)
    ops.upsample_fwd(srcd, False, out=big[..., :c])
    assert torch.equal(big[..., c:].cpu(), torch.full((n, 2 * h, 2 * w, width - c), 7.0)), "upsample: neighbours touched"
    close(big[..., :c], REF.upsample_fwd(srcc, False), 2e-6, "upsample into slice")
    if c >= 16:      # tensor-core convolution writing a ragged slice
        xc, xd = _poisoned(ops, n, h, w, 32, 92)
        wc, wd = weight_pair(ops, c, 32, 3, 93)
        full = fresh()
        ops.conv_fprop(xd, wd, None, ACT_NONE, 0.0, out=full[..., :c])
        assert torch.equal(full[..., c:].cpu(), torch.full((n, h, w, width - c), 7.0)), "conv: neighbours touched"
        close(full[..., :c], REF.conv_fprop(xc, wc, None), 4e-3, "conv into slice")


@pytest.mark.parametrize("hw", [(256, 448), (37, 53), (11, 11), (9, 40)])
def test_ssim_kernel_against_pytorch_msssim_restatement(cuda_ops, hw):
    """mi_ssim_accumulate (utils.py:195-204 -> pytorch_msssim/__init__.py:19-75) vs the ATen restatement: structured
    and noisy pairs, windows clipped by small images, ragged tiles."""
    from meta_interpolation_b200.utils import gaussian_1d
    g = torch.Generator().manual_seed(4)
    h, w = hw
    base = torch.rand(3, h, w, generator=g)
    for noise in (0.02, 0.5):
        other = (base + noise * torch.randn(3, h, w, generator=g)).clamp(-0.2, 1.2)
        win = gaussian_1d(min(11, h, w))
        sr, sd = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64, device="cuda")
        REF.ssim_accumulate(base, other, win, sr)
        cuda_ops.ssim_accumulate(base.cuda(), other.cuda(), win, sd)
        n = 3 * (h - win.numel() + 1) * (w - win.numel() + 1)
        assert abs(sr.item() - sd.item()) / n <= 2e-5, (hw, noise, sr.item() / n, sd.item() / n)


def test_round_tf32_kernel_and_rounded_weight_copies(cuda_ops):
    """mi_round_tf32 == cvt.rna.tf32.f32 (restated with integer arithmetic in oracle/ops_ref.py): flat buffers,
    NHWC rows with a pixel stride (neighbouring channels untouched), in place and out of place; the rotated dgrad
    copy is rounded on request."""
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4096, generator=g) * 3
    assert torch.equal(cuda_ops.round_tf32(x.cuda().clone()).cpu(), REF.round_tf32(x.clone()))
    buf = torch.randn(2, 5, 7, 12, generator=g)
    view_c, view_r = buf.cuda().clone(), buf.clone()
    cuda_ops.round_tf32(view_c[..., 2:9])                 # 7 channels of a 12-float pixel stride
    REF.round_tf32(view_r[..., 2:9])
    assert torch.equal(view_c.cpu(), view_r)
    assert torch.equal(view_c.cpu()[..., :2], buf[..., :2]) and torch.equal(view_c.cpu()[..., 9:], buf[..., 9:])
    w = cuda_ops.empty_weight(40, 37, 3)
    w.copy_(torch.randn(40, 3, 3, 37, generator=g))
    wt = cuda_ops.weight_to_dgrad(w, rnd=True)
    assert torch.equal(wt.cpu(), REF.weight_to_dgrad(w.cpu(), rnd=True))
    assert not torch.equal(wt.cpu(), REF.weight_to_dgrad(w.cpu(), rnd=False))


def test_tensor_core_conv_outputs_lie_on_the_tf32_grid(cuda_ops):
    """TF32 operand convention: fprop / dgrad epilogues of the tensor-core engine store round-to-nearest TF32 values,
    so the next conv's hardware truncation is exact."""
    if not cuda_ops.tf32_rn:
        pytest.skip("exact-fp32 engine forced")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = cuda_ops.empty_act(2, 40, 56, 64); x.copy_(torch.rand(2, 40, 56, 64, device="cuda", generator=g))
    w = cuda_ops.empty_weight(51, 64, 3); w.copy_(torch.rand(51, 3, 3, 64, device="cuda", generator=g) - 0.5)
    b = torch.rand(51, device="cuda", generator=g)
    y = cuda_ops.conv_fprop(x, w, b, 1, 0.0)
    assert torch.equal(y.contiguous(), cuda_ops.round_tf32(y.contiguous().clone()))
    dx = cuda_ops.conv_dgrad(y, w)
    assert torch.equal(dx.contiguous(), cuda_ops.round_tf32(dx.contiguous().clone()))
