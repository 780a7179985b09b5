"""Oracle: functional backbone forwards on plain ATen CPU ops (test infrastructure).

Each function takes NCHW fp32 frames and a ``{name: tensor}`` dict keyed exactly
like the reference's ``named_parameters()`` and returns
what the reference's ``net.forward(frame0, frame1, params=...)`` returns.

``routed`` / ``meta`` split: the reference only substitutes fast weights in
part of each net; everything else reads the
stored meta-parameters.  ``fast`` holds the fast weights, ``meta`` the stored
parameters; a key missing from ``fast`` falls back to ``meta`` exactly where the
reference would have used ``self.<module>`` without params.
"""
import torch
import torch.nn.functional as F

from .sepconv_op import FunctionSepconvCPU


# --------------------------------------------------------------------------- sepconv

SEPCONV_ROUTED_PREFIXES = ("moduleConv1.", "moduleConv2.", "moduleConv3.", "moduleConv4.", "moduleConv5.",
                           "moduleDeconv5.", "moduleDeconv4.", "moduleDeconv3.", "moduleDeconv2.")


def sepconv_is_routed(name):
    """True if the reference feeds this tensor from ``params`` (sepconv/model.py:276-309);
    moduleUpsample2-5 and the four Subnets never are (:292,297,302,307,346-347)."""
    return name.startswith(SEPCONV_ROUTED_PREFIXES)


def _conv(x, p, prefix):
    return F.conv2d(x, p[prefix + ".weight"], p[prefix + ".bias"], stride=1, padding=1)


def _basic(x, p, prefix):
    # sepconv/model.py:172-181: (conv3x3 + ReLU) x3 at indices 0,2,4
    for i in (0, 2, 4):
        x = F.relu(_conv(x, p, "%s.%d" % (prefix, i)))
    return x


def _up2_align(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def _upsample_block(x, p, prefix):
    # sepconv/model.py:211-215: bilinear x2 (align_corners=True) -> conv -> ReLU
    return F.relu(_conv(_up2_align(x), p, prefix + ".1"))


def _subnet(x, p, prefix):
    # sepconv/model.py:184-194
    x = F.relu(_conv(x, p, prefix + ".0"))
    x = F.relu(_conv(x, p, prefix + ".2"))
    x = F.relu(_conv(x, p, prefix + ".4"))
    x = _up2_align(x)
    return _conv(x, p, prefix + ".7")


def sepconv_geometry(height, width):
    """Canvas geometry of sepconv/model.py:252-266: returns (canvas_h, canvas_w)."""
    pw = 25 + width + 25
    ph = 25 + height + 25
    if pw != ((pw >> 7) << 7):
        pw = ((pw >> 7) + 1) << 7
    if ph != ((ph >> 7) << 7):
        ph = ((ph >> 7) + 1) << 7
    return ph, pw


def sepconv_forward(frame0, frame1, fast, meta, sepconv_fn=None):
    """reference sepconv/model.py:252-350 (MetaNetwork.forward)."""
    sepconv_fn = sepconv_fn or FunctionSepconvCPU.apply
    height, width = frame0.shape[2], frame0.shape[3]
    ch, cw = sepconv_geometry(height, width)
    pad_in = [25, cw - 25 - width, 25, ch - 25 - height]
    f0 = F.pad(frame0, pad_in, mode="replicate")
    f1 = F.pad(frame1, pad_in, mode="replicate")
    x = torch.cat([f0, f1], 1)

    p = dict(meta)
    p.update({k: v for k, v in fast.items() if sepconv_is_routed(k)})
    m = meta  # un-routed modules always read the stored parameters

    c1 = _basic(x, p, "moduleConv1")
    c2 = _basic(F.avg_pool2d(c1, 2, 2), p, "moduleConv2")
    c3 = _basic(F.avg_pool2d(c2, 2, 2), p, "moduleConv3")
    c4 = _basic(F.avg_pool2d(c3, 2, 2), p, "moduleConv4")
    c5 = _basic(F.avg_pool2d(c4, 2, 2), p, "moduleConv5")
    d5 = _basic(F.avg_pool2d(c5, 2, 2), p, "moduleDeconv5")
    comb = _upsample_block(d5, m, "moduleUpsample5") + c5
    d4 = _basic(comb, p, "moduleDeconv4")
    comb = _upsample_block(d4, m, "moduleUpsample4") + c4
    d3 = _basic(comb, p, "moduleDeconv3")
    comb = _upsample_block(d3, m, "moduleUpsample3") + c3
    d2 = _basic(comb, p, "moduleDeconv2")
    comb = _upsample_block(d2, m, "moduleUpsample2") + c2

    pad25 = lambda t: F.pad(t, [25, 25, 25, 25], mode="replicate")
    dot1 = sepconv_fn(pad25(f0), _subnet(comb, m, "moduleVertical1"), _subnet(comb, m, "moduleHorizontal1"))
    dot2 = sepconv_fn(pad25(f1), _subnet(comb, m, "moduleVertical2"), _subnet(comb, m, "moduleHorizontal2"))
    out = dot1 + dot2
    return out[:, :, 25:25 + height, 25:25 + width]


def sepconv_param_shapes():
    """Ordered (name, shape) list == reference named_parameters() (SURVEY Appx H)."""
    out = []

    def conv(name, cin, cout):
        out.append((name + ".weight", (cout, cin, 3, 3)))
        out.append((name + ".bias", (cout,)))

    def basic(name, cin, cout):
        conv(name + ".0", cin, cout)
        conv(name + ".2", cout, cout)
        conv(name + ".4", cout, cout)

    basic("moduleConv1", 6, 32)
    basic("moduleConv2", 32, 64)
    basic("moduleConv3", 64, 128)
    basic("moduleConv4", 128, 256)
    basic("moduleConv5", 256, 512)
    basic("moduleDeconv5", 512, 512)
    conv("moduleUpsample5.1", 512, 512)
    basic("moduleDeconv4", 512, 256)
    conv("moduleUpsample4.1", 256, 256)
    basic("moduleDeconv3", 256, 128)
    conv("moduleUpsample3.1", 128, 128)
    basic("moduleDeconv2", 128, 64)
    conv("moduleUpsample2.1", 64, 64)
    for sub in ("moduleVertical1", "moduleVertical2", "moduleHorizontal1", "moduleHorizontal2"):
        conv(sub + ".0", 64, 64)
        conv(sub + ".2", 64, 64)
        conv(sub + ".4", 64, 51)
        conv(sub + ".7", 51, 51)
    return out


from . import backbones_flow as _bf  # noqa: E402

BACKBONES = {
    "sepconv": dict(forward=sepconv_forward, is_routed=sepconv_is_routed, shapes=sepconv_param_shapes),
    "voxelflow": dict(forward=_bf.voxelflow_forward, is_routed=_bf.voxelflow_is_routed,
                      shapes=_bf.voxelflow_param_shapes),
    "superslomo": dict(forward=_bf.superslomo_forward, is_routed=_bf.superslomo_is_routed,
                       shapes=_bf.superslomo_param_shapes),
    "rrin": dict(forward=_bf.rrin_forward, is_routed=_bf.rrin_is_routed, shapes=_bf.rrin_param_shapes),
    "cain": dict(forward=_bf.cain_forward, is_routed=_bf.cain_is_routed, shapes=_bf.cain_param_shapes),
}

SUPERSLOMO_MEAN = (0.429, 0.431, 0.397)


def denormalise(model, x):
    """Prediction / target back to [0,1] (meta_learning_system.py:70-73,78-79,434-447)."""
    if model == "superslomo":
        return x + torch.tensor(SUPERSLOMO_MEAN, dtype=x.dtype).view(-1, 1, 1)
    if model == "voxelflow":
        return (x * 127.5 + 127.5) / 255.0
    return x


def set_torch_seed(seed):
    """meta_learning_system.py:16-26: numpy RandomState(seed).randint -> torch.manual_seed."""
    import numpy as np
    rng = np.random.RandomState(seed=seed)
    torch.manual_seed(seed=int(rng.randint(0, 999999)))
    return rng


def seeded_params(model, seed=12345):
    """Default init of the reference (xavier-uniform weights, zero biases,
    model_utils.py:329-333) drawn in construction order under the reference's
    seeding (meta_learning_system.py:48), so the same seed gives the same
    tensors as ``SceneAdaptiveInterpolation(args).net`` without the reference."""
    from collections import OrderedDict
    set_torch_seed(seed)
    if model == "voxelflow":
        return _bf.voxelflow_seeded_params()
    out = OrderedDict()
    for name, shape in BACKBONES[model]["shapes"]():
        if name.endswith(".weight") and len(shape) == 4:
            w = torch.empty(*shape)
            torch.nn.init.xavier_uniform_(w)
            out[name] = w
        else:
            out[name] = torch.zeros(*shape)
    return out
