"""SepConv plugin (drop-in for the reference's ``sepconv/model.py:MetaNetwork``).

Same constructor, parameter names/shapes and ``forward(tensorFirst, tensorSecond,
params=None, **kwargs)`` contract as reference sepconv/model.py:168-375; the
compute is a tape of sm_100a kernels on an NHWC canvas:

* ``modulePaddingInput`` (replicate pad to a multiple of 128 incl. the 25-px
  border, :254-269) is folded into the kernel that builds the 6-channel canvas;
* ``modulePad`` (+25 replicate, :244-245) and ``modulePaddingOutput`` (crop, :264-266)
  are folded into the separable-convolution kernel, which reads the raw frame with
  clamped coordinates and only evaluates the H x W window that survives the crop
  (SURVEY.md Appendix A.1);
* only moduleConv1-5 / moduleDeconv5-2 take fast weights; moduleUpsample2-5 and the
  four Subnets always use the stored parameters (:292,297,302,307,346-347; quirk Q1).
"""
import torch

from ..backbone import MetaBackbone
from ..ops import ACT_NONE, ACT_RELU

_ROUTED = ("moduleConv1.", "moduleConv2.", "moduleConv3.", "moduleConv4.", "moduleConv5.",
           "moduleDeconv5.", "moduleDeconv4.", "moduleDeconv3.", "moduleDeconv2.")


def canvas_size(height, width, border=25):
    """reference sepconv/model.py:254-261: pad by the border, then up to a multiple of 128."""
    pw, ph = border + width + border, border + height + border
    if pw != ((pw >> 7) << 7):
        pw = ((pw >> 7) + 1) << 7
    if ph != ((ph >> 7) << 7):
        ph = ((ph >> 7) + 1) << 7
    return ph, pw


class MetaNetwork(MetaBackbone):
    FILTER = 51

    def __init__(self, resume=False, strModel='lf', ops=None):
        super().__init__(ops)
        self.padding = [25, 25, 25, 25]

        def init(name, shape):
            if name.endswith(".weight"):
                w = torch.empty(*shape)
                torch.nn.init.xavier_uniform_(w)   # reference model_utils.py:329-330
                return w
            return torch.zeros(*shape)             # :333

        self._build_parameters(init)
        if resume:
            print('Loading model: pretrained_models/sepconv_base_%s.pth' % strModel)
            self.load_state_dict(torch.load('pretrained_models/sepconv_base_' + strModel + '.pth'))

    # ------------------------------------------------------------------ structure
    def conv_specs(self):
        specs = []

        def basic(name, cin, cout):
            specs.append((name + ".0", cin, cout, 3, True))
            specs.append((name + ".2", cout, cout, 3, True))
            specs.append((name + ".4", cout, cout, 3, True))

        basic("moduleConv1", 6, 32)
        basic("moduleConv2", 32, 64)
        basic("moduleConv3", 64, 128)
        basic("moduleConv4", 128, 256)
        basic("moduleConv5", 256, 512)
        basic("moduleDeconv5", 512, 512)
        specs.append(("moduleUpsample5.1", 512, 512, 3, True))
        basic("moduleDeconv4", 512, 256)
        specs.append(("moduleUpsample4.1", 256, 256, 3, True))
        basic("moduleDeconv3", 256, 128)
        specs.append(("moduleUpsample3.1", 128, 128, 3, True))
        basic("moduleDeconv2", 128, 64)
        specs.append(("moduleUpsample2.1", 64, 64, 3, True))
        for sub in ("moduleVertical1", "moduleVertical2", "moduleHorizontal1", "moduleHorizontal2"):
            specs.append((sub + ".0", 64, 64, 3, True))
            specs.append((sub + ".2", 64, 64, 3, True))
            specs.append((sub + ".4", 64, 51, 3, True))
            specs.append((sub + ".7", 51, 51, 3, True))
        return specs

    def is_routed(self, param_name):
        return param_name.startswith(_ROUTED)

    # ------------------------------------------------------------------ graph
    def _basic(self, t, x, name):
        for i in (0, 2, 4):
            x = t.conv(x, "%s.%d" % (name, i), ACT_RELU)
        return x

    def _up(self, t, x, name):
        return t.conv(t.upsample(x, True), name + ".1", ACT_RELU)

    def _subnet(self, t, x, name):
        x = t.conv(x, name + ".0", ACT_RELU)
        x = t.conv(x, name + ".2", ACT_RELU)
        x = t.conv(x, name + ".4", ACT_RELU)
        x = t.upsample(x, True)
        return t.conv(x, name + ".7", ACT_NONE)

    def build_graph(self, t, frame0, frame1):
        """reference sepconv/model.py:252-350; frames are NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        from ..tape import Var
        n, _, height, width = frame0.shape
        ch, cw = canvas_size(height, width)
        canvas = Var(t.ops.frames_to_canvas(frame0, frame1, ch, cw, 25, 25, 0), requires_grad=False)

        c1 = self._basic(t, canvas, "moduleConv1")
        c2 = self._basic(t, t.avgpool(c1), "moduleConv2")
        c3 = self._basic(t, t.avgpool(c2), "moduleConv3")
        c4 = self._basic(t, t.avgpool(c3), "moduleConv4")
        c5 = self._basic(t, t.avgpool(c4), "moduleConv5")
        d5 = self._basic(t, t.avgpool(c5), "moduleDeconv5")
        comb = t.add(self._up(t, d5, "moduleUpsample5"), c5)
        d4 = self._basic(t, comb, "moduleDeconv4")
        comb = t.add(self._up(t, d4, "moduleUpsample4"), c4)
        d3 = self._basic(t, comb, "moduleDeconv3")
        comb = t.add(self._up(t, d3, "moduleUpsample3"), c3)
        d2 = self._basic(t, comb, "moduleDeconv2")
        comb = t.add(self._up(t, d2, "moduleUpsample2"), c2)

        v1 = self._subnet(t, comb, "moduleVertical1")
        h1 = self._subnet(t, comb, "moduleHorizontal1")
        v2 = self._subnet(t, comb, "moduleVertical2")
        h2 = self._subnet(t, comb, "moduleHorizontal2")
        # output pixel (i,j) = canvas pixel (i+25, j+25); its 51x51 window starts at frame (i-25, j-25)
        dot1 = t.sepconv(frame0, v1, h1, height, width, 25, 25, -25, -25)
        dot2 = t.sepconv(frame1, v2, h2, height, width, 25, 25, -25, -25)
        return t.add_nchw(dot1, dot2)
