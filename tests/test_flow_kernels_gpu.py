"""Glue / head kernels of the flow-based and attention backbones (glue.cu, heads.cu) against oracle/ops_ref.py."""
import pytest
import torch

from oracle.ops_ref import RefOps
from meta_interpolation_b200.ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, WG_ACCUM, WG_STORE)
from test_kernels_gpu import act_pair, close

pytestmark = pytest.mark.gpu
REF = RefOps()


def vec_pair(c, seed, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    cpu = torch.rand(c, generator=g) * (hi - lo) + lo
    return cpu, cpu.cuda()


@pytest.mark.parametrize("c", [3, 16, 64, 130])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_RELU])
def test_bn_eval_fwd_bwd(cuda_ops, c, act):
    """Frozen (eval-mode) batch norm of VoxelFlow (voxel_flow.py:241-263, 352-355) fused with its ReLU."""
    n, h, w = 2, 10, 12
    xc, xd = act_pair(cuda_ops, n, h, w, c, 1)
    gam, bet = vec_pair(c, 2, 0.5, 1.5), vec_pair(c, 3)
    mean, var = vec_pair(c, 4, -0.2, 0.2), vec_pair(c, 5, 0.5, 2.0)
    yc = REF.bn_eval_fwd(xc, gam[0], bet[0], mean[0], var[0], 1e-5, act, 0.0)
    yd = cuda_ops.bn_eval_fwd(xd, gam[1], bet[1], mean[1], var[1], 1e-5, act, 0.0)
    close(yd, yc, 2e-6, "bn fwd")
    gyc, gyd = act_pair(cuda_ops, n, h, w, c, 6)
    for mode, acc_dx in ((WG_STORE, False), (WG_ACCUM, True)):
        dxc, dxd = act_pair(cuda_ops, n, h, w, c, 7)
        dgc, dbc = vec_pair(c, 8), vec_pair(c, 9)
        dg_c, dg_d, db_c, db_d = dgc[0].clone(), dgc[1].clone(), dbc[0].clone(), dbc[1].clone()
        REF.bn_eval_bwd(gyc, yc, xc, gam[0], mean[0], var[0], 1e-5, act, 0.0, dxc, acc_dx, dg_c, db_c, mode, 0.5)
        cuda_ops.bn_eval_bwd(gyd, yd, xd, gam[1], mean[1], var[1], 1e-5, act, 0.0, dxd, acc_dx, dg_d, db_d, mode, 0.5)
        close(dxd, dxc, 2e-6, "bn dx")
        close(dg_d, dg_c, 2e-5, "bn dgamma mode %d" % mode)
        close(db_d, db_c, 2e-5, "bn dbeta mode %d" % mode)
    # dx is optional (first layer of the encoder: the input is data)
    dg_c, dg_d, db_c, db_d = (torch.zeros(c), torch.zeros(c, device="cuda"), torch.zeros(c),
                              torch.zeros(c, device="cuda"))
    REF.bn_eval_bwd(gyc, yc, xc, gam[0], mean[0], var[0], 1e-5, act, 0.0, None, False, dg_c, db_c, WG_STORE, 1.0)
    cuda_ops.bn_eval_bwd(gyd, yd, xd, gam[1], mean[1], var[1], 1e-5, act, 0.0, None, False, dg_d, db_d, WG_STORE, 1.0)
    close(dg_d, dg_c, 2e-5, "bn dgamma (no dx)")


@pytest.mark.parametrize("op", [0, 1, 2, 3])
@pytest.mark.parametrize("c,cb", [(3, 3), (3, 1), (8, 8), (5, 1), (1, 1)])
def test_binary_broadcast_fwd_bwd(cuda_ops, op, c, cb):
    n, h, w = 2, 7, 9
    ac, ad = act_pair(cuda_ops, n, h, w, c, 10)
    bc, bd = act_pair(cuda_ops, n, h, w, cb, 11)
    if op == 3:     # keep the divisor away from zero
        bc.copy_(bc.abs() + 0.5)
        bd.copy_(bd.abs() + 0.5)
    close(cuda_ops.binary_fwd(op, ad, bd), REF.binary_fwd(op, ac, bc), 2e-6, "binary fwd")
    goc, god = act_pair(cuda_ops, n, h, w, c, 12)
    for acc in (False, True):
        gac, gad = act_pair(cuda_ops, n, h, w, c, 13)
        gbc, gbd = act_pair(cuda_ops, n, h, w, cb, 14)
        REF.binary_bwd(op, ac, bc, goc, gac, acc, gbc, acc)
        cuda_ops.binary_bwd(op, ad, bd, god, gad, acc, gbd, acc)
        close(gad, gac, 2e-6, "binary ga")
        close(gbd, gbc, 1e-5, "binary gb")
    gac, gad = act_pair(cuda_ops, n, h, w, c, 15)
    REF.binary_bwd(op, ac, bc, goc, gac, False, None, False)
    cuda_ops.binary_bwd(op, ad, bd, god, gad, False, None, False)
    close(gad, gac, 2e-6, "binary ga only")


@pytest.mark.parametrize("c", [2, 3, 32])
def test_affine_act_clamp(cuda_ops, c):
    n, h, w = 2, 6, 10
    xc, xd = act_pair(cuda_ops, n, h, w, c, 20)
    close(cuda_ops.affine(xd, 0.5, 0.25), REF.affine(xc, 0.5, 0.25), 1e-6, "affine")
    oc, od = act_pair(cuda_ops, n, h, w, c, 21)
    REF.affine(xc, -2.0, 0.0, out=oc, accumulate=True)
    cuda_ops.affine(xd, -2.0, 0.0, out=od, accumulate=True)
    close(od, oc, 1e-6, "affine accumulate")
    for act in (ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH):
        close(cuda_ops.act_fwd(xd, act, 0.2), REF.act_fwd(xc, act, 0.2), 2e-6, "act_fwd %d" % act)
    close(cuda_ops.clamp_fwd(xd, -0.3, 0.4), REF.clamp_fwd(xc, -0.3, 0.4), 0.0, "clamp")
    gc, gd = act_pair(cuda_ops, n, h, w, c, 22)
    for acc in (False, True):
        dc, dd = act_pair(cuda_ops, n, h, w, c, 23)
        REF.clamp_bwd(gc, xc, dc, -0.3, 0.4, acc)
        cuda_ops.clamp_bwd(gd, xd, dd, -0.3, 0.4, acc)
        close(dd, dc, 1e-6, "clamp bwd")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_blend_fwd_bwd(cuda_ops, mode):
    """Visibility blend of SuperSloMo (superslomo/model.py:619-633), mask blends of RRIN / VoxelFlow."""
    n, h, w, c = 2, 9, 11, 3
    ac, ad = act_pair(cuda_ops, n, h, w, c, 30)
    bc, bd = act_pair(cuda_ops, n, h, w, c, 31)
    m0c, m0d = act_pair(cuda_ops, n, h, w, 1, 32)
    m1c, m1d = act_pair(cuda_ops, n, h, w, 1, 33)
    for t in (m0c, m0d, m1c, m1d):      # masks live in (0,1)
        t.copy_(t * 0.4 + 0.5)
    m1 = (m1c, m1d) if mode == 0 else (None, None)
    w0, w1, eps = 0.5, 0.5, (1e-8 if mode == 1 else 0.0)
    yc = REF.blend_fwd(ac, bc, m0c, m1[0], w0, w1, eps, mode)
    yd = cuda_ops.blend_fwd(ad, bd, m0d, m1[1], w0, w1, eps, mode)
    close(yd, yc, 3e-6, "blend fwd")
    goc, god = act_pair(cuda_ops, n, h, w, c, 34)
    for acc in (False, True):
        outs_c = [act_pair(cuda_ops, n, h, w, cc, 35 + i) for i, cc in enumerate((c, c, 1, 1))]
        gc = [p[0] for p in outs_c]
        gd = [p[1] for p in outs_c]
        if mode != 0:
            gc[3] = gd[3] = None
        REF.blend_bwd(ac, bc, m0c, m1[0], goc, gc[0], gc[1], gc[2], gc[3], acc, w0, w1, eps, mode)
        cuda_ops.blend_bwd(ad, bd, m0d, m1[1], god, gd[0], gd[1], gd[2], gd[3], acc, w0, w1, eps, mode)
        for i, nm in enumerate(("ga", "gb", "gm0", "gm1")):
            if gc[i] is not None:
                close(gd[i], gc[i], 1e-5, "blend %s mode %d acc %s" % (nm, mode, acc))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("geom", [(1, 6, 7, 5), (2, 10, 12, 192), (1, 5, 5, 8), (1, 5, 9, 64), (2, 18, 19, 192)])
def test_ring_fix_and_fold(cuda_ops, mode, geom):
    """1-pixel ring of CAIN's ReflectionPad2d(1) convs (model_utils.py:825-827) kept in place; fold = its adjoint."""
    n, h, w, c = geom
    xc, xd = act_pair(cuda_ops, n, h, w, c, 40)
    REF.ring_fix(xc, mode)
    cuda_ops.ring_fix(xd, mode)
    close(xd, xc, 0.0, "ring_fix")
    gc, gd = act_pair(cuda_ops, n, h, w, c, 41)
    REF.ring_fold(gc, mode)
    cuda_ops.ring_fold(gd, mode)
    close(gd, gc, 1e-6, "ring_fold")


def test_ring_rejects_degenerate_interior(cuda_ops):
    """An interior narrower than 3 pixels cannot be reflected without self-overlap: bad-argument, not garbage."""
    from meta_interpolation_b200._lib import MiB200Error
    x = cuda_ops.empty_act(1, 4, 9, 8)
    with pytest.raises(MiB200Error):
        cuda_ops.ring_fix(x, 1)
    with pytest.raises(MiB200Error):
        cuda_ops.ring_fold(x, 0)


@pytest.mark.parametrize("hw,r", [((16, 24), 8), ((13, 21), 8), ((8, 8), 4), ((64, 40), 8)])
def test_space_depth_roundtrip_and_reference(cuda_ops, hw, r):
    """CAIN's sub-mean + InOutPaddings + pixel_shuffle(1/8) (model_utils.py:11-28, 202-217; cain/model.py:70-94)."""
    n, (h, w) = 2, hw
    mult = 2 * r
    ph, pw = (mult - h % mult) % mult, (mult - w % mult) % mult
    pad_top, pad_left = ph // 2, pw // 2
    oh, ow = (h + ph) // r, (w + pw) // r
    if pad_top >= h or ph - pad_top >= h or pad_left >= w or pw - pad_left >= w:
        pytest.skip("reflection pad wider than the frame")
    g = torch.Generator().manual_seed(50)
    f0, f1 = torch.rand(n, 3, h, w, generator=g), torch.rand(n, 3, h, w, generator=g)
    m0c, m1c = REF.channel_mean_nchw(f0), REF.channel_mean_nchw(f1)
    m0d, m1d = cuda_ops.channel_mean_nchw(f0.cuda()), cuda_ops.channel_mean_nchw(f1.cuda())
    close(m0d, m0c, 2e-6, "channel mean")
    sc = REF.space_to_depth(f0, f1, m0c, m1c, pad_top, pad_left, oh, ow, r)
    sd = cuda_ops.space_to_depth(f0.cuda(), f1.cuda(), m0d, m1d, pad_top, pad_left, oh, ow, r)
    close(sd, sc, 2e-6, "space_to_depth")
    # the decoder's tensor: 3*r*r channels on the ringed grid
    xc, xd = act_pair(cuda_ops, n, oh + 2, ow + 2, 3 * r * r, 51)
    yc = REF.depth_to_space(xc, m0c, m1c, h, w, pad_top, pad_left, r)
    yd = cuda_ops.depth_to_space(xd, m0d, m1d, h, w, pad_top, pad_left, r)
    close(yd, yc, 2e-6, "depth_to_space")
    go = torch.rand(n, 3, h, w, generator=g)
    gic, gid = act_pair(cuda_ops, n, oh + 2, ow + 2, 3 * r * r, 52)
    REF.depth_to_space_bwd(go, gic, pad_top, pad_left, r)
    cuda_ops.depth_to_space_bwd(go.cuda(), gid, pad_top, pad_left, r)
    close(gid, gic, 0.0, "depth_to_space_bwd")
    # size-independent property: <depth_to_space(x) - mean, g> == <x, depth_to_space_bwd(g)>
    lhs = ((yd - 0.5 * (m0d + m1d).view(n, 3, 1, 1)) * go.cuda()).double().sum().item()
    rhs = (xd * gid).double().sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


@pytest.mark.parametrize("geom", [(1, 9, 9, 32, 1), (2, 33, 20, 8, 0), (1, 12, 12, 64, 1)])
def test_interior_reduce_scalar_and_vector_paths(cuda_ops, geom):
    """16-byte-aligned rows take the float4 kernel; a channel slice that starts one float into a wider buffer (rows not
    16-byte aligned) takes the scalar one.  Same answer."""
    n, h, w, c, ring = geom
    xc, xd = act_pair(cuda_ops, n, h, w, c + 4, 64)
    mc, md = act_pair(cuda_ops, n, h, w, c + 4, 65)
    for lo in (0, 1):
        xs, ms = xd[..., lo:lo + c], md[..., lo:lo + c]
        xr, mr = xc[..., lo:lo + c].contiguous(), mc[..., lo:lo + c].contiguous()
        close(cuda_ops.interior_reduce(xs, None, ring, 0.25), REF.interior_reduce(xr, None, ring, 0.25), 2e-5, "reduce")
        close(cuda_ops.interior_reduce(xs, ms, ring, 1.0), REF.interior_reduce(xr, mr, ring, 1.0), 2e-5, "reduce * mul")


@pytest.mark.parametrize("geom", [(1, 10, 10, 192, 1), (2, 6, 9, 12, 1), (2, 5, 7, 64, 0), (1, 66, 66, 192, 1),
                                  (2, 130, 130, 192, 1), (1, 40, 56, 100, 1)])   # many pixels per cluster rank
def test_channel_attention_pieces(cuda_ops, geom):
    """Global average pool, per-channel rescale + residual and their adjoints (MetaCALayer, model_utils.py:931-953)."""
    n, h, w, c, ring = geom
    xc, xd = act_pair(cuda_ops, n, h, w, c, 60)
    mc, md = act_pair(cuda_ops, n, h, w, c, 61)
    scale = 1.0 / ((h - 2 * ring) * (w - 2 * ring))
    pc, pd = REF.interior_reduce(xc, None, ring, scale), cuda_ops.interior_reduce(xd, None, ring, scale)
    close(pd, pc, 2e-5, "interior_reduce")
    close(cuda_ops.interior_reduce(xd, md, ring, 1.0), REF.interior_reduce(xc, mc, ring, 1.0), 2e-5,
          "interior_reduce with multiplier")
    sc, sd = act_pair(cuda_ops, n, 1, 1, c, 62)
    rc, rd = act_pair(cuda_ops, n, h, w, c, 63)
    close(cuda_ops.scale_add(xd, sd, rd), REF.scale_add(xc, sc, rc), 2e-6, "scale_add")
    close(cuda_ops.scale_add(xd, sd, None), REF.scale_add(xc, sc, None), 2e-6, "scale (no residual)")
    for acc in (False, True):
        dc, dd = act_pair(cuda_ops, n, h, w, c, 64)
        REF.scale_bwd(xc, sc, dc, acc)
        cuda_ops.scale_bwd(xd, sd, dd, acc)
        close(dd, dc, 2e-6, "scale_bwd")
    dc, dd = act_pair(cuda_ops, n, h, w, c, 65)
    REF.interior_bcast_add(sc, dc, ring, scale)
    cuda_ops.interior_bcast_add(sd, dd, ring, scale)
    close(dd, dc, 2e-6, "interior_bcast_add")


@pytest.mark.parametrize("variant,sx,sy", [(0, 1.0, 1.0), (0, -1.0, -1.0), (1, 1.0, 1.0), (1, -1.0, -1.0)])
def test_warp_constant_image_property_at_baseline_size(cuda_ops, variant, sx, sy):
    """Backward warp at the BASELINE frame size (256x448): bilinear weights sum to one, so a constant image warps to
    itself wherever the sample stays inside (variant 0, zeros padding) / everywhere (variant 1, border clamp), and
    the flow gradient of a constant image is zero there."""
    n, h, w, c = 2, 256, 448, 3
    img = cuda_ops.empty_act(n, h, w, c)
    img.fill_(0.75)
    g = torch.Generator().manual_seed(70)
    flow = cuda_ops.empty_act(n, h, w, 2)
    flow.copy_(((torch.rand(n, h, w, 2, generator=g) - 0.5) * (6.0 if variant == 0 else 0.04)).cuda())
    out = cuda_ops.warp_fwd(img, flow, variant, sx, sy)
    region = out[:, 8:-8, 8:-8, :] if variant == 0 else out
    assert (region - 0.75).abs().max().item() <= 1e-5
    go = cuda_ops.empty_act(n, h, w, c)
    go.copy_(torch.rand(n, h, w, c, generator=g).cuda())
    gf = cuda_ops.empty_act(n, h, w, 2)
    cuda_ops.warp_bwd(img, flow, go, gf, variant, sx, sy)
    region = gf[:, 8:-8, 8:-8, :]
    assert region.abs().max().item() <= 1e-4
