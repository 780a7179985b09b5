"""CAIN plugin (drop-in for the reference's ``cain/model.py:MetaCAIN``).

Same constructor, parameter names/shapes (``encoder.interpolate.*``, 494 tensors) and
``forward(x1, x2, params=None, **kwargs)`` contract as reference cain/model.py:53-118; building blocks
``MetaInterpolation / MetaResidualGroup / MetaRCAB / MetaCALayer / MetaConvNorm`` of model_utils.py:821-1053.

* ``sub_mean`` + reflection padding to a multiple of 128 + ``PixelShuffle(1/8)`` of both frames (model_utils.py:
  11-28, 202-217) are ONE kernel writing the 384-channel NHWC feature map; ``PixelShuffle(8)`` + crop + mean
  shift (cain/model.py:84-94) are one kernel back to NCHW;
* every feature map lives in the interior of a buffer with a one-pixel ring: ``ReflectionPad2d(1)`` + conv
  (MetaConvNorm) becomes "fill the ring in place, run the zero-padding tcgen05 conv over the whole buffer"; the
  head / tail convs (zero padding) use a zeroed ring.  No padded copy of any activation is ever made;
* channel attention = interior mean kernel, the two 1x1 convs on [n,1,1,C] tensors through the ordinary conv ABI
  (so their inner-loop update is fused like every other conv), and one rescale+residual kernel.
The reference walks ~1000 Python-level dict peelings per forward (35 ms); here names are resolved once.
"""
import torch

from ..backbone import MetaBackbone
from ..ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID
from ..padding import reflect_pads, xavier_or_zero

GROUPS, BLOCKS, FEATS, REDUCTION = 5, 12, 192, 16
_PRE = "encoder.interpolate."


class MetaCAIN(MetaBackbone):
    def __init__(self, depth=3, resume=False, ops=None):
        super().__init__(ops)
        self.depth = depth
        assert 3 * (4 ** depth) == FEATS
        self._build_parameters(xavier_or_zero)
        if resume:
            print('Loading model: pretrained_models/cain_base.pth')
            checkpoint = torch.load('pretrained_models/cain_base.pth')
            self.load_state_dict({k.replace("module.", ""): v for k, v in checkpoint['state_dict'].items()})

    # ------------------------------------------------------------------ structure
    def conv_specs(self):
        c = FEATS
        specs = [(_PRE + "headConv", 2 * c, c, 3, True)]
        for g in range(GROUPS):
            for b in range(BLOCKS):
                base = _PRE + "body.%d.body.%d.body." % (g, b)
                specs.append((base + "0.conv", c, c, 3, True))
                specs.append((base + "2.conv", c, c, 3, True))
                specs.append((base + "3.conv_du.0", c, c // REDUCTION, 1, True))
                specs.append((base + "3.conv_du.2", c // REDUCTION, c, 1, True))
            specs.append((_PRE + "body.%d.body.%d.conv" % (g, BLOCKS), c, c, 3, True))
        specs.append((_PRE + "tailConv", c, c, 3, True))
        return specs

    # ------------------------------------------------------------------ graph
    def build_graph(self, t, frame0, frame1):
        """reference MetaCAIN.forward :70-94; frames NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        ops = t.ops
        n, _, height, width = frame0.shape
        left, right, top, bottom = reflect_pads(height, width, 7)
        r = 1 << self.depth
        oh, ow = (height + top + bottom) // r, (width + left + right) // r
        m0, m1 = ops.channel_mean_nchw(frame0), ops.channel_mean_nchw(frame1)
        feats = t.data(ops.space_to_depth(frame0, frame1, m0, m1, top, left, oh, ow, r))
        zero, refl = ops.RING_ZERO, ops.RING_REFLECT

        x = t.ring_conv(feats, _PRE + "headConv", ACT_NONE, 0.0, zero)
        res = x
        for g in range(GROUPS):
            group_in = res
            for b in range(BLOCKS):                                   # MetaRCAB, model_utils.py:957-990
                base = _PRE + "body.%d.body.%d.body." % (g, b)
                o = t.ring_conv(res, base + "0.conv", ACT_LEAKY, 0.2, refl)
                o = t.ring_conv(o, base + "2.conv", ACT_NONE, 0.0, refl)
                y = t.interior_mean(o, 1)                             # MetaCALayer, :931-955
                y = t.conv(y, base + "3.conv_du.0", ACT_RELU)
                s = t.conv(y, base + "3.conv_du.2", ACT_SIGMOID)
                res = t.scale_add(o, s, res, 1)
            tail = t.ring_conv(res, _PRE + "body.%d.body.%d.conv" % (g, BLOCKS), ACT_NONE, 0.0, refl)
            res = t.add(tail, group_in)                               # MetaResidualGroup, :994-1011
        res = t.add(res, x)                                           # MetaInterpolation, :1045-1047
        y = t.ring_conv(res, _PRE + "tailConv", ACT_NONE, 0.0, zero)
        return t.depth_to_space(y, m0, m1, height, width, top, left, r)
