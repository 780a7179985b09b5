"""Oracle: functional forwards of the flow / attention backbones on plain ATen CPU ops (test infrastructure).

Same contract as ``oracle/backbones.py``: NCHW fp32 frames and ``{name: tensor}`` dicts keyed like the
reference's ``named_parameters()``; ``fast`` holds the fast weights, ``meta`` the
stored parameters, and a tensor the reference never routes (Appendix A, Q2/Q2b) is always read from ``meta``.

* voxelflow  : voxelflow/core/models/voxel_flow.py:357-509 (MetaVoxelFlow.forward), meshgrid :9-17
* superslomo : superslomo/model.py:565-645 (MetaSuperSloMo.forward), MetaUNet :457-544, down/up :11-153,
               backWarp :231-303, getFlowCoeff/getWarpCoeff :310-379
* rrin       : rrin/model.py:74-130 (MetaRRIN.process/forward), warp :8-21, rrin/unet.py:96-208
* cain       : cain/model.py:70-94 (MetaCAIN.forward), model_utils.py:11-28,202-217,821-1053
"""
import torch
import torch.nn.functional as F


def _pick(fast, meta, routed):
    p = dict(meta)
    p.update({k: v for k, v in fast.items() if routed(k)})
    return p


def _conv(x, p, name, pad):
    return F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=1, padding=pad)


def reflect_pads(height, width, shift):
    """Pad amounts (left, right, top, bottom) up to a multiple of 2**shift, split floor/ceil
    (voxel_flow.py:360-368, superslomo/model.py:567-575 shift 6; model_utils.py:17-28 shift 7)."""
    pw = ph = 0
    if width != ((width >> shift) << shift):
        pw = (((width >> shift) + 1) << shift) - width
    if height != ((height >> shift) << shift):
        ph = (((height >> shift) + 1) << shift) - height
    return pw // 2, pw - pw // 2, ph // 2, ph - ph // 2


def _pad_in(x, pads):
    return F.pad(x, list(pads), mode="reflect") if any(pads) else x


def _crop_out(x, pads):
    l, r, t, b = pads
    return x[:, :, t:x.shape[2] - b, l:x.shape[3] - r]


# --------------------------------------------------------------------------- voxelflow

def voxelflow_is_routed(name):
    """conv weights (and conv4.bias) come from ``params``; the BN scale/shift are in the dict but
    ``self.convN_bn(x)`` ignores it (voxel_flow.py:379,385,... )."""
    return "_bn." not in name


def voxelflow_param_shapes():
    out = []
    for name, cin, cout, k in (("conv1", 6, 64, 5), ("conv2", 64, 128, 5), ("conv3", 128, 256, 3),
                               ("bottleneck", 256, 256, 3), ("deconv1", 512, 256, 3), ("deconv2", 384, 128, 5),
                               ("deconv3", 192, 64, 5)):
        out.append((name + ".weight", (cout, cin, k, k)))
        out.append((name + "_bn.weight", (cout,)))
        out.append((name + "_bn.bias", (cout,)))
    out.append(("conv4.weight", (3, 64, 5, 5)))
    out.append(("conv4.bias", (3,)))
    return out


def voxelflow_seeded_params():
    """MetaVoxelFlow.__init__ (voxel_flow.py:241-274): every MetaConv2dLayer draws a xavier-uniform weight at
    construction, then all conv weights are redrawn N(0, 0.01) in module order; BN scale 1 / shift 0."""
    from collections import OrderedDict
    shapes = voxelflow_param_shapes()
    for name, shape in shapes:
        if len(shape) == 4:
            torch.nn.init.xavier_uniform_(torch.empty(*shape))   # consumed and overwritten below
    out = OrderedDict()
    for name, shape in shapes:
        if len(shape) == 4:
            out[name] = torch.empty(*shape).normal_(0, 0.01)
        elif name.endswith("_bn.weight"):
            out[name] = torch.ones(*shape)
        else:
            out[name] = torch.zeros(*shape)
    return out


def _bn_eval(x, p, name, eps=1e-5):
    # frozen BatchNorm2d with the constructor's running statistics (mean 0, var 1): voxel_flow.py:352-355
    c = x.shape[1]
    return F.batch_norm(x, torch.zeros(c), torch.ones(c), p[name + ".weight"], p[name + ".bias"], False, 0.0, eps)


def _linspace_grid(h, w):
    gx = torch.linspace(-1.0, 1.0, w).view(1, 1, w).expand(1, h, w)
    gy = torch.linspace(-1.0, 1.0, h).view(1, h, 1).expand(1, h, w)
    return gx, gy


def voxelflow_forward(frame0, frame1, fast, meta):
    p = _pick(fast, meta, voxelflow_is_routed)
    pads = reflect_pads(frame0.shape[2], frame0.shape[3], 6)
    inp = _pad_in(torch.cat([frame0, frame1], 1), pads)
    h, w = inp.shape[2], inp.shape[3]
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)

    def block(t, name, pad):
        return F.relu(_bn_eval(_conv(t, p, name, pad), p, name + "_bn"))

    c1 = block(inp, "conv1", 2)
    c2 = block(F.max_pool2d(c1, 2, 2), "conv2", 2)
    c3 = block(F.max_pool2d(c2, 2, 2), "conv3", 1)
    x = block(F.max_pool2d(c3, 2, 2), "bottleneck", 1)
    x = block(torch.cat([up(x), c3], 1), "deconv1", 1)
    x = block(torch.cat([up(x), c2], 1), "deconv2", 2)
    x = block(torch.cat([up(x), c1], 1), "deconv3", 2)
    x = torch.tanh(_conv(x, p, "conv4", 2))

    flow = 0.5 * x[:, 0:2]
    mask = x[:, 2:3]
    gx, gy = _linspace_grid(h, w)
    n = inp.shape[0]
    gx, gy = gx.repeat(n, 1, 1), gy.repeat(n, 1, 1)
    grid1 = torch.stack([gx - flow[:, 0], gy - flow[:, 1]], dim=3)
    grid2 = torch.stack([gx + flow[:, 0], gy + flow[:, 1]], dim=3)
    o1 = F.grid_sample(inp[:, 0:3], grid1, padding_mode="border", align_corners=True)
    o2 = F.grid_sample(inp[:, 3:6], grid2, padding_mode="border", align_corners=True)
    m = (0.5 * (1.0 + mask)).repeat(1, 3, 1, 1)
    return _crop_out(m * o1 + (1.0 - m) * o2, pads)


# --------------------------------------------------------------------------- shared by superslomo / rrin

def backwarp(img, flow):
    """superslomo/model.py:292-302 == rrin/model.py:8-21: grid 2*((x+u)/W-0.5), bilinear, zeros,
    align_corners=False  (sample position x+u-0.5 in pixel units)."""
    _, _, h, w = img.shape
    gx = torch.arange(w, dtype=torch.float32).view(1, 1, w).expand(1, h, w)
    gy = torch.arange(h, dtype=torch.float32).view(1, h, 1).expand(1, h, w)
    x = gx + flow[:, 0]
    y = gy + flow[:, 1]
    grid = torch.stack((2 * (x / w - 0.5), 2 * (y / h - 0.5)), dim=3)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


# --------------------------------------------------------------------------- superslomo

SLOMO_T = [0.125 + 0.125 * i for i in range(7)]   # np.linspace(0.125, 0.875, 7), superslomo/model.py:308


def superslomo_is_routed(name):
    return True


def _slomo_unet_shapes(prefix, cin, cout):
    out = []

    def conv(name, ci, co, k):
        out.append((prefix + name + ".weight", (co, ci, k, k)))
        out.append((prefix + name + ".bias", (co,)))

    conv("conv1", cin, 32, 7)
    conv("conv2", 32, 32, 7)
    for i, (ci, co, k) in enumerate(((32, 64, 5), (64, 128, 3), (128, 256, 3), (256, 512, 3), (512, 512, 3)), 1):
        conv("down%d.conv1" % i, ci, co, k)
        conv("down%d.conv2" % i, co, co, k)
    for i, (ci, co) in enumerate(((512, 512), (512, 256), (256, 128), (128, 64), (64, 32)), 1):
        conv("up%d.conv1" % i, ci, co, 3)
        conv("up%d.conv2" % i, 2 * co, co, 3)
    conv("conv3", 32, cout, 3)
    return out


def superslomo_param_shapes():
    return _slomo_unet_shapes("flowComp.", 6, 4) + _slomo_unet_shapes("arbTimeFlowIntrp.", 20, 5)


def _slomo_unet(x, p, prefix):
    lr = lambda t: F.leaky_relu(t, negative_slope=0.1)
    x = lr(_conv(x, p, prefix + "conv1", 3))
    s1 = lr(_conv(x, p, prefix + "conv2", 3))
    skips = [s1]
    x = s1
    for i, k in enumerate((5, 3, 3, 3, 3), 1):
        x = F.avg_pool2d(x, 2)
        x = lr(_conv(x, p, prefix + "down%d.conv1" % i, k // 2))
        x = lr(_conv(x, p, prefix + "down%d.conv2" % i, k // 2))
        skips.append(x)
    skips.pop()     # down5's output is the bottom of the U, not a skip
    for i in range(1, 6):
        x = F.interpolate(x, scale_factor=2, mode="bilinear")     # align_corners default False (model.py:139)
        x = lr(_conv(x, p, prefix + "up%d.conv1" % i, 1))
        x = lr(_conv(torch.cat((x, skips.pop()), 1), p, prefix + "up%d.conv2" % i, 1))
    return lr(_conv(x, p, prefix + "conv3", 1))


def superslomo_forward(frame0, frame1, fast, meta, ind=3, full=False):
    p = _pick(fast, meta, superslomo_is_routed)
    pads = reflect_pads(frame0.shape[2], frame0.shape[3], 6)
    i0, i1 = _pad_in(frame0, pads), _pad_in(frame1, pads)
    t = SLOMO_T[ind]
    flow = _slomo_unet(torch.cat((i0, i1), 1), p, "flowComp.")
    f01, f10 = flow[:, :2], flow[:, 2:]
    c00 = c11 = -(1 - t) * t
    c01, c10 = t * t, (1 - t) * (1 - t)
    ft0 = c00 * f01 + c01 * f10
    ft1 = c10 * f01 + c11 * f10
    g0 = backwarp(i0, ft0)
    g1 = backwarp(i1, ft1)
    intrp = _slomo_unet(torch.cat((i0, i1, f01, f10, ft1, ft0, g1, g0), 1), p, "arbTimeFlowIntrp.")
    ft0f = intrp[:, :2] + ft0
    ft1f = intrp[:, 2:4] + ft1
    v0 = torch.sigmoid(intrp[:, 4:5])
    v1 = 1 - v0
    g0f = backwarp(i0, ft0f)
    g1f = backwarp(i1, ft1f)
    w0, w1 = 1 - t, t
    out = (w0 * v0 * g0f + w1 * v1 * g1f) / (w0 * v0 + w1 * v1)
    out = _crop_out(out, pads)
    if not full:
        return out
    aux = {"bidirectional_flow": (_crop_out(f01, pads), _crop_out(f10, pads)),
           "warped_intermediate_frames": (_crop_out(g0, pads), _crop_out(g1, pads)),
           "warped_input_frames": (_crop_out(backwarp(i0, f10), pads), _crop_out(backwarp(i1, f01), pads))}
    return out, aux


# --------------------------------------------------------------------------- rrin

def rrin_is_routed(name):
    """``self.Mask(temp)`` is called outside the ``if params`` branch (rrin/model.py:100-101): the Mask U-Net
    always uses the stored parameters."""
    return not name.startswith("Mask.")


def _rrin_unet_shapes(prefix, cin, cout, depth, wf=5):
    out = []

    def conv(name, ci, co):
        out.append((prefix + name + ".weight", (co, ci, 3, 3)))
        out.append((prefix + name + ".bias", (co,)))

    prev = cin
    for i in range(depth):
        c = 2 ** (wf + i)
        conv("down_path.%d.block.0" % i, prev, c)
        conv("down_path.%d.block.2" % i, c, c)
        prev = c
    conv("midconv", prev, prev)
    for j, i in enumerate(reversed(range(depth - 1))):
        c = 2 ** (wf + i)
        conv("up_path.%d.up.1" % j, prev, c)
        conv("up_path.%d.conv_block.block.0" % j, prev, c)
        conv("up_path.%d.conv_block.block.2" % j, c, c)
        prev = c
    conv("last", prev, cout)
    return out


RRIN_UNETS = (("Mask.", 16, 2, 4), ("Flow_L.", 6, 4, 5), ("refine_flow.", 10, 4, 4), ("final.", 9, 3, 4))


def rrin_param_shapes():
    out = []
    for prefix, cin, cout, depth in RRIN_UNETS:
        out += _rrin_unet_shapes(prefix, cin, cout, depth)
    return out


def _rrin_unet(x, p, prefix, depth):
    lr = lambda t: F.leaky_relu(t, negative_slope=0.1)
    blocks = []
    for i in range(depth):
        x = lr(_conv(x, p, prefix + "down_path.%d.block.0" % i, 1))
        x = lr(_conv(x, p, prefix + "down_path.%d.block.2" % i, 1))
        if i != depth - 1:
            blocks.append(x)
            x = F.avg_pool2d(x, 2)
    x = lr(_conv(x, p, prefix + "midconv", 1))
    for j in range(depth - 1):
        up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        up = _conv(up, p, prefix + "up_path.%d.up.1" % j, 1)
        bridge = blocks[-j - 1]
        dy = (bridge.shape[2] - up.shape[2]) // 2
        dx = (bridge.shape[3] - up.shape[3]) // 2
        bridge = bridge[:, :, dy:dy + up.shape[2], dx:dx + up.shape[3]]
        x = torch.cat((up, bridge), 1)
        x = lr(_conv(x, p, prefix + "up_path.%d.conv_block.block.0" % j, 1))
        x = lr(_conv(x, p, prefix + "up_path.%d.conv_block.block.2" % j, 1))
    return _conv(x, p, prefix + "last", 1)


def rrin_forward(frame0, frame1, fast, meta, t=0.5):
    p = _pick(fast, meta, rrin_is_routed)
    pads = reflect_pads(frame0.shape[2], frame0.shape[3], 7)
    x0, x1 = _pad_in(frame0, pads), _pad_in(frame1, pads)
    x = torch.cat((x0, x1), 1)
    flow = _rrin_unet(x, p, "Flow_L.", 5)
    f01, f10 = flow[:, :2], flow[:, 2:4]
    ft0 = -(1 - t) * t * f01 + t * t * f10
    ft1 = (1 - t) * (1 - t) * f01 - t * (1 - t) * f10
    ref = _rrin_unet(torch.cat((ft0, ft1, x), 1), p, "refine_flow.", 4)
    ft0 = ft0 + ref[:, :2]
    ft1 = ft1 + ref[:, 2:4]
    xt1 = backwarp(x0, ft0)
    xt2 = backwarp(x1, ft1)
    mask = torch.sigmoid(_rrin_unet(torch.cat((ft0, ft1, x, xt1, xt2), 1), p, "Mask.", 4))
    w1, w2 = (1 - t) * mask[:, 0:1], t * mask[:, 1:2]
    out = (w1 * xt1 + w2 * xt2) / (w1 + w2 + 1e-8)
    final = _rrin_unet(torch.cat((x0, x1, out), 1), p, "final.", 4) + out
    return _crop_out(final.clamp(0, 1), pads)


# --------------------------------------------------------------------------- cain

CAIN_GROUPS, CAIN_BLOCKS, CAIN_FEATS, CAIN_REDUCTION = 5, 12, 192, 16


def cain_is_routed(name):
    return True


def cain_param_shapes():
    out = []
    pre = "encoder.interpolate."

    def conv(name, ci, co, k):
        out.append((pre + name + ".weight", (co, ci, k, k)))
        out.append((pre + name + ".bias", (co,)))

    c = CAIN_FEATS
    conv("headConv", 2 * c, c, 3)
    for g in range(CAIN_GROUPS):
        for b in range(CAIN_BLOCKS):
            base = "body.%d.body.%d.body." % (g, b)
            conv(base + "0.conv", c, c, 3)
            conv(base + "2.conv", c, c, 3)
            conv(base + "3.conv_du.0", c, c // CAIN_REDUCTION, 1)
            conv(base + "3.conv_du.2", c // CAIN_REDUCTION, c, 1)
        conv("body.%d.body.%d.conv" % (g, CAIN_BLOCKS), c, c, 3)
    conv("tailConv", c, c, 3)
    return out


def _space_to_depth(x, r):
    # model_utils.py:202-217 with scale_factor = 1/r
    n, c, h, w = x.shape
    v = x.contiguous().view(n, c, h // r, r, w // r, r)
    return v.permute(0, 1, 3, 5, 2, 4).contiguous().view(n, c * r * r, h // r, w // r)


def _depth_to_space(x, r):
    # model_utils.py:202-217 with scale_factor = r
    n, c, h, w = x.shape
    oc = c // (r * r)
    v = x.contiguous().view(n, oc, r, r, h, w)
    return v.permute(0, 1, 4, 2, 5, 3).contiguous().view(n, oc, h * r, w * r)


def _conv_reflect(x, p, name):
    # MetaConvNorm: ReflectionPad2d(1) + conv(padding=0)  (model_utils.py:821-849)
    return F.conv2d(F.pad(x, [1, 1, 1, 1], mode="reflect"), p[name + ".weight"], p[name + ".bias"])


def cain_forward(frame0, frame1, fast, meta, depth=3):
    p = _pick(fast, meta, cain_is_routed)
    pre = "encoder.interpolate."
    m1 = frame0.mean(2, keepdim=True).mean(3, keepdim=True)
    m2 = frame1.mean(2, keepdim=True).mean(3, keepdim=True)
    pads = reflect_pads(frame0.shape[2], frame0.shape[3], 7)
    x1, x2 = _pad_in(frame0 - m1, pads), _pad_in(frame1 - m2, pads)
    r = 2 ** depth
    x = _conv(torch.cat([_space_to_depth(x1, r), _space_to_depth(x2, r)], 1), p, pre + "headConv", 1)
    res = x
    for g in range(CAIN_GROUPS):
        gin = res
        for b in range(CAIN_BLOCKS):
            base = pre + "body.%d.body.%d.body." % (g, b)
            o = _conv_reflect(res, p, base + "0.conv")
            o = F.leaky_relu(o, 0.2)
            o = _conv_reflect(o, p, base + "2.conv")
            y = o.mean(dim=(2, 3), keepdim=True)
            y = F.relu(_conv(y, p, base + "3.conv_du.0", 0))
            y = torch.sigmoid(_conv(y, p, base + "3.conv_du.2", 0))
            res = o * y + res
        res = _conv_reflect(res, p, pre + "body.%d.body.%d.conv" % (g, CAIN_BLOCKS)) + gin
    res = res + x
    out = _depth_to_space(_conv(res, p, pre + "tailConv", 1), r)
    return _crop_out(out, pads) + (m1 + m2) / 2
