"""SuperSloMo plugin (drop-in for the reference's ``superslomo/model.py:MetaSuperSloMo``).

Same constructor, parameter names/shapes (``flowComp.*`` / ``arbTimeFlowIntrp.*``) and
``forward(I0, I1, ind=3, params=None, **kwargs) -> (Ft_p, {...})`` contract as reference
superslomo/model.py:547-645; the compute is a tape of sm_100a kernels on an NHWC canvas:

* the reflection input padding to a multiple of 64 (:567-578) is folded into the kernel that builds the 6-channel
  canvas, the output crop (:632-635) into the NHWC->NCHW window kernel;
* every U-Net level writes its skip connection and its up-conv straight into the pre-allocated concat buffer of
  the matching ``up`` block (no ``torch.cat`` copies, :139-152); the 20-channel input of ``arbTimeFlowIntrp``
  (:606) is assembled the same way -- ``flowComp``'s last conv, the interpolated flows and the two warps write
  their channel slices in place;
* ``backWarp`` (:231-303) is the zero-padded bilinear gather kernel sampling at (x+u-0.5, y+v-0.5);
  the visibility-weighted fusion (:617-630) is one fused blend kernel forward and one backward.
All 92 tensors are routed from ``params``, so nothing is skipped in support passes.
"""
import torch

from ..backbone import MetaBackbone
from ..ops import ACT_LEAKY, ACT_SIGMOID
from ..padding import reflect_pads, xavier_or_zero

T_GRID = [0.125 + 0.125 * i for i in range(7)]     # np.linspace(0.125, 0.875, 7), reference :308
_DOWN = ((32, 64, 5), (64, 128, 3), (128, 256, 3), (256, 512, 3), (512, 512, 3))
_UP = ((512, 512), (512, 256), (256, 128), (128, 64), (64, 32))
_CH = (32, 64, 128, 256, 512)


class MetaSuperSloMo(MetaBackbone):
    def __init__(self, device=None, resume=False, ops=None):
        super().__init__(ops)
        self.device = device
        self.backwarp = None
        self._ind = 3
        self._aux = None
        self._build_parameters(xavier_or_zero)
        if resume:
            print('Loading model: pretrained_models/superslomo_base.pth')
            checkpoint = torch.load('pretrained_models/superslomo_base.pth')
            sd = {'flowComp.' + k: v for k, v in checkpoint['state_dictFC'].items()}
            sd.update({'arbTimeFlowIntrp.' + k: v for k, v in checkpoint['state_dictAT'].items()})
            self.load_state_dict(sd)

    # ------------------------------------------------------------------ structure
    def conv_specs(self):
        specs = []
        for prefix, cin, cout in (("flowComp.", 6, 4), ("arbTimeFlowIntrp.", 20, 5)):
            specs.append((prefix + "conv1", cin, 32, 7, True))
            specs.append((prefix + "conv2", 32, 32, 7, True))
            for i, (ci, co, k) in enumerate(_DOWN, 1):
                specs.append((prefix + "down%d.conv1" % i, ci, co, k, True))
                specs.append((prefix + "down%d.conv2" % i, co, co, k, True))
            for i, (ci, co) in enumerate(_UP, 1):
                specs.append((prefix + "up%d.conv1" % i, ci, co, 3, True))
                specs.append((prefix + "up%d.conv2" % i, 2 * co, co, 3, True))
            specs.append((prefix + "conv3", 32, cout, 3, True))
        return specs

    # ------------------------------------------------------------------ graph
    def _unet(self, t, x, pre, out=None):
        """reference MetaUNet.forward :495-544; ``down`` :69-78, ``up`` :139-152."""
        n, h, w, _ = x.data.shape
        cats = [t.concat_buffer(n, h >> i, w >> i, 2 * _CH[i]) for i in range(5)]
        x = t.conv(x, pre + "conv1", ACT_LEAKY, 0.1)
        x = t.conv(x, pre + "conv2", ACT_LEAKY, 0.1, out=cats[0][..., 32:64])
        skips = [x]
        for i in range(1, 6):
            x = t.avgpool(x)
            x = t.conv(x, pre + "down%d.conv1" % i, ACT_LEAKY, 0.1)
            if i < 5:
                c = _CH[i]
                x = t.conv(x, pre + "down%d.conv2" % i, ACT_LEAKY, 0.1, out=cats[i][..., c:2 * c])
                skips.append(x)
            else:
                x = t.conv(x, pre + "down5.conv2", ACT_LEAKY, 0.1)
        for i in range(1, 6):
            lvl = 5 - i
            c = _CH[lvl]
            x = t.upsample(x, False)
            u = t.conv(x, pre + "up%d.conv1" % i, ACT_LEAKY, 0.1, out=cats[lvl][..., 0:c])
            cat = t.as_var_of_slices(cats[lvl], [(u, 0, c), (skips[lvl], c, 2 * c)])
            x = t.conv(cat, pre + "up%d.conv2" % i, ACT_LEAKY, 0.1)
        return t.conv(x, pre + "conv3", ACT_LEAKY, 0.1, out=out)

    def build_graph(self, t, frame0, frame1):
        """reference MetaSuperSloMo.forward :565-645; frames NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        ops = t.ops
        n, _, height, width = frame0.shape
        left, right, top, bottom = reflect_pads(height, width, 6)
        ch, cw = height + top + bottom, width + left + right
        tt = T_GRID[self._ind]
        canvas = ops.frames_to_canvas(frame0, frame1, ch, cw, top, left, 1)
        i0, i1 = canvas[..., 0:3], canvas[..., 3:6]

        # arbTimeFlowIntrp input: (I0, I1, F_0_1, F_1_0, F_t_1, F_t_0, g_I1_F_t_1, g_I0_F_t_0), :606
        cat20 = t.concat_buffer(n, ch, cw, 20)
        flow = self._unet(t, t.data(canvas), "flowComp.", out=cat20[..., 6:10])
        f01, f10 = t.slice(flow, 0, 2), t.slice(flow, 2, 4)
        c00 = c11 = -(1 - tt) * tt            # getFlowCoeff :310-343
        c01, c10 = tt * tt, (1 - tt) * (1 - tt)
        ft0 = t.lincomb([(c00, f01), (c01, f10)], out=cat20[..., 12:14])
        ft1 = t.lincomb([(c10, f01), (c11, f10)], out=cat20[..., 10:12])
        g0 = t.warp(i0, ft0, 0, out=cat20[..., 17:20])
        g1 = t.warp(i1, ft1, 0, out=cat20[..., 14:17])
        x = t.concat(cat20, [(flow, 6, 10), (ft1, 10, 12), (ft0, 12, 14), (g1, 14, 17), (g0, 17, 20)],
                     consts=[(canvas, 0, 6)])
        intrp = self._unet(t, x, "arbTimeFlowIntrp.")

        ft0f = t.lincomb([(1.0, t.slice(intrp, 0, 2)), (1.0, ft0)])
        ft1f = t.lincomb([(1.0, t.slice(intrp, 2, 4)), (1.0, ft1)])
        v0 = t.act(t.slice(intrp, 4, 5), ACT_SIGMOID)
        g0f = t.warp(i0, ft0f, 0)
        g1f = t.warp(i1, ft1f, 0)
        # (C0 V0 g0 + C1 V1 g1) / (C0 V0 + C1 V1) with V1 = 1 - V0, getWarpCoeff :346-379
        out = t.blend(g0f, g1f, v0, None, 1 - tt, tt, 0.0, ops.BLEND_RATIO_COMPLEMENT)
        self._aux = dict(canvas=canvas, f01=f01.data, f10=f10.data, g0=g0.data, g1=g1.data,
                         window=(top, left, height, width))
        # the same tensors as tape Vars: the `Super` loss (loss.py:258-274) differentiates through them
        self.aux_vars = dict(f01=f01, f10=f10, g0=g0, g1=g1, i0=i0, i1=i1, window=(top, left, height, width))
        return t.to_nchw(out, top, left, height, width)

    # ------------------------------------------------------------------ reference plugin API
    def forward(self, I0, I1, ind=3, params=None, **kwargs):
        self._ind = int(ind)
        try:
            out = super().forward(I0, I1, params=params)
        finally:
            self._ind = 3
        aux, ops = self._aux, self.ops
        top, left, h, w = aux["window"]
        crop = lambda t: ops.nhwc_window_to_nchw(t, top, left, h, w)
        i0, i1 = aux["canvas"][..., 0:3], aux["canvas"][..., 3:6]
        # the three tuples feed only the `Super` loss (loss.py:258-274), which is outside this path; they are
        # returned as constants
        extras = {'bidirectional_flow': (crop(aux["f01"]), crop(aux["f10"])),
                  'warped_intermediate_frames': (crop(aux["g0"]), crop(aux["g1"])),
                  'warped_input_frames': (crop(ops.warp_fwd(i0, aux["f10"], 0)), crop(ops.warp_fwd(i1, aux["f01"], 0)))}
        self._aux = None
        return out, extras
