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


def _up2_sources(o, in_size):
    """(i0, i1) read by output index ``o`` of the x2 align_corners=True upsample (fp32 arithmetic of the kernel)."""
    import numpy as np
    out_size = 2 * in_size
    scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
    i0 = min(int(scale * np.float32(o)), in_size - 1)
    return i0, i0 + (1 if i0 < in_size - 1 else 0)


def _axis_roi(win0, win_len, full_hi, convs_low=3):
    """One axis of the Subnet region of interest.  The separable convolution reads its filters on canvas positions
    [win0, win0+win_len) only (the crop of sepconv/model.py:350).  Walking the Subnet backwards: the final 3x3 conv
    needs one more position on each side (hi window); its x2 upsample reads low-resolution sources i0..i1 of that
    window; the three 3x3 convs below need ``convs_low`` more on each side (low window).  Windows are clipped at the
    canvas, where the zero padding of the full evaluation applies unchanged.  Returns (hi0, hi_len, lo0, lo_len)."""
    full_lo = full_hi // 2
    hi0, hi1 = max(win0 - 1, 0), min(win0 + win_len + 1, full_hi)            # [hi0, hi1)
    src0, src1 = _up2_sources(hi0, full_lo)[0], _up2_sources(hi1 - 1, full_lo)[1]
    lo0, lo1 = max(src0 - convs_low, 0), min(src1 + convs_low + 1, full_lo)
    for o in range(hi0, hi1):       # every source of the hi window lies inside the low window (no clamping)
        a, b = _up2_sources(o, full_lo)
        assert lo0 <= a and b < lo1
    return hi0, hi1 - hi0, lo0, lo1 - lo0


class MetaNetwork(MetaBackbone):
    FILTER = 51
    SUBNET_ROI = True   # evaluate the four Subnets only where the cropped prediction can see them

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

    def _subnet(self, t, x, name, roi=None):
        x = t.conv(x, name + ".0", ACT_RELU)
        x = t.conv(x, name + ".2", ACT_RELU)
        x = t.conv(x, name + ".4", ACT_RELU)
        if roi is None:
            x = t.upsample(x, True)
        else:
            x = t.upsample_window(x, True, roi["full_lo"], roi["lo_origin"], roi["hi_origin"], roi["hi_hw"])
        return t.conv(x, name + ".7", ACT_NONE)

    def build_graph(self, t, frame0, frame1):
        """reference sepconv/model.py:252-350; frames are NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        from ..tape import Var
        n, _, height, width = frame0.shape
        ch, cw = canvas_size(height, width)
        canvas = Var(t.ops.frames_to_canvas(frame0, frame1, ch, cw, 25, 25, 0, rnd=t.ops.tf32_rn), requires_grad=False)
        canvas.clean = t.ops.tf32_rn

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

        # Region of interest: the prediction is cropped to canvas rows [25, 25+H) x cols [25, 25+W)
        # (modulePaddingOutput, reference :264-266, :350), so the filters outside that window are dead values.  The
        # Subnets (52 % of the backbone's FLOPs) run on the crop of `comb` whose receptive field covers the window;
        # inside the window every value equals the full-canvas evaluation (3 pixels of slack absorb the wrong zero
        # padding at the crop edges), and gradients vanish identically outside it.
        roi, gy0, gx0 = None, 25, 25
        if self.SUBNET_ROI:
            hy0, hh, ly0, lh = _axis_roi(25, height, ch)
            hx0, hw, lx0, lw = _axis_roi(25, width, cw)
            if lh * lw < (ch // 2) * (cw // 2):
                roi = dict(full_lo=(ch // 2, cw // 2), lo_origin=(ly0, lx0), hi_origin=(hy0, hx0), hi_hw=(hh, hw))
                comb = t.crop(comb, ly0, lx0, lh, lw)
                gy0, gx0 = 25 - hy0, 25 - hx0
        v1 = self._subnet(t, comb, "moduleVertical1", roi)
        h1 = self._subnet(t, comb, "moduleHorizontal1", roi)
        v2 = self._subnet(t, comb, "moduleVertical2", roi)
        h2 = self._subnet(t, comb, "moduleHorizontal2", roi)
        # output pixel (i,j) = canvas pixel (i+25, j+25); its 51x51 window starts at frame (i-25, j-25)
        dot1 = t.sepconv(frame0, v1, h1, height, width, gy0, gx0, -25, -25)
        dot2 = t.sepconv(frame1, v2, h2, height, width, gy0, gx0, -25, -25)
        return t.add_nchw(dot1, dot2)
