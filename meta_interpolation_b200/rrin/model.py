"""RRIN plugin (drop-in for the reference's ``rrin/model.py:MetaRRIN``).

Same constructor, parameter names/shapes (``Mask.* / Flow_L.* / refine_flow.* / final.*``)
and ``forward(input0, input1, t=0.5, params=None, **kwargs)`` contract as reference rrin/model.py:61-151;
the four U-Nets (rrin/unet.py:96-208) run as a tape of sm_100a kernels on an NHWC canvas padded by reflection
to a multiple of 128 (model_utils.py:17-28).  Bridges and up-convs are written straight into the concat buffers;
``warp`` (:8-21) is the zero-padded bilinear gather kernel, the mask-weighted fusion (:101-103) one blend kernel.
The ``Mask`` U-Net is never fed from ``params`` (:100-101): it always reads the stored parameters and
its weight gradients are skipped in support passes.
"""
import torch

from ..backbone import MetaBackbone
from ..ops import ACT_LEAKY, ACT_NONE, ACT_SIGMOID
from ..padding import reflect_pads, xavier_or_zero

_UNETS = (("Mask.", 16, 2, 4), ("Flow_L.", 6, 4, 5), ("refine_flow.", 10, 4, 4), ("final.", 9, 3, 4))


class MetaRRIN(MetaBackbone):
    def __init__(self, level=3, resume=False, ops=None):
        super().__init__(ops)
        self._t = 0.5
        self._build_parameters(xavier_or_zero)
        if resume:
            print('Loading model: pretrained_models/rrin_base.pth')
            self.load_state_dict(torch.load('pretrained_models/rrin_base.pth'))

    # ------------------------------------------------------------------ structure
    def conv_specs(self):
        specs = []
        for prefix, cin, cout, depth in _UNETS:
            prev = cin
            for i in range(depth):
                c = 32 << i
                specs.append((prefix + "down_path.%d.block.0" % i, prev, c, 3, True))
                specs.append((prefix + "down_path.%d.block.2" % i, c, c, 3, True))
                prev = c
            specs.append((prefix + "midconv", prev, prev, 3, True))
            for j, i in enumerate(reversed(range(depth - 1))):
                c = 32 << i
                specs.append((prefix + "up_path.%d.up.1" % j, prev, c, 3, True))
                specs.append((prefix + "up_path.%d.conv_block.block.0" % j, prev, c, 3, True))
                specs.append((prefix + "up_path.%d.conv_block.block.2" % j, c, c, 3, True))
                prev = c
            specs.append((prefix + "last", prev, cout, 3, True))
        return specs

    def is_routed(self, param_name):
        return not param_name.startswith("Mask.")

    # ------------------------------------------------------------------ graph
    def _unet(self, t, x, pre, depth, final_act=ACT_NONE, out=None):
        """reference rrin/unet.py MetaUNet.forward :125-152, conv block :154-170, up block :173-208."""
        n, h, w, _ = x.data.shape
        assert h % (1 << (depth - 1)) == 0 and w % (1 << (depth - 1)) == 0   # center_crop is the identity
        cats = [t.concat_buffer(n, h >> i, w >> i, 2 * (32 << i)) for i in range(depth - 1)]
        blocks = []
        for i in range(depth):
            c = 32 << i
            x = t.conv(x, pre + "down_path.%d.block.0" % i, ACT_LEAKY, 0.1)
            if i != depth - 1:
                x = t.conv(x, pre + "down_path.%d.block.2" % i, ACT_LEAKY, 0.1, out=cats[i][..., c:2 * c])
                blocks.append(x)
                x = t.avgpool(x)
            else:
                x = t.conv(x, pre + "down_path.%d.block.2" % i, ACT_LEAKY, 0.1)
        x = t.conv(x, pre + "midconv", ACT_LEAKY, 0.1)
        for j in range(depth - 1):
            lvl = depth - 2 - j
            c = 32 << lvl
            x = t.upsample(x, False)
            u = t.conv(x, pre + "up_path.%d.up.1" % j, ACT_NONE, out=cats[lvl][..., 0:c])
            cat = t.as_var_of_slices(cats[lvl], [(u, 0, c), (blocks[lvl], c, 2 * c)])
            x = t.conv(cat, pre + "up_path.%d.conv_block.block.0" % j, ACT_LEAKY, 0.1)
            x = t.conv(x, pre + "up_path.%d.conv_block.block.2" % j, ACT_LEAKY, 0.1)
        return t.conv(x, pre + "last", final_act, out=out)

    def build_graph(self, t, frame0, frame1):
        """reference MetaRRIN.process/forward :74-130; frames NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        ops = t.ops
        n, _, height, width = frame0.shape
        left, right, top, bottom = reflect_pads(height, width, 7)
        ch, cw = height + top + bottom, width + left + right
        tt = self._t
        canvas = ops.frames_to_canvas(frame0, frame1, ch, cw, top, left, 1)
        x0, x1 = canvas[..., 0:3], canvas[..., 3:6]

        flow = self._unet(t, t.data(canvas), "Flow_L.", 5)
        f01, f10 = t.slice(flow, 0, 2), t.slice(flow, 2, 4)
        cat10 = t.concat_buffer(n, ch, cw, 10)                       # (Flow_t_0, Flow_t_1, x), :86
        ft0 = t.lincomb([(-(1 - tt) * tt, f01), (tt * tt, f10)], out=cat10[..., 0:2])
        ft1 = t.lincomb([((1 - tt) * (1 - tt), f01), (-tt * (1 - tt), f10)], out=cat10[..., 2:4])
        ref = self._unet(t, t.concat(cat10, [(ft0, 0, 2), (ft1, 2, 4)], consts=[(canvas, 4, 10)]), "refine_flow.", 4)

        cat16 = t.concat_buffer(n, ch, cw, 16)                       # (Flow_t_0, Flow_t_1, x, xt1, xt2), :99
        ft0 = t.lincomb([(1.0, ft0), (1.0, t.slice(ref, 0, 2))], out=cat16[..., 0:2])
        ft1 = t.lincomb([(1.0, ft1), (1.0, t.slice(ref, 2, 4))], out=cat16[..., 2:4])
        xt1 = t.warp(x0, ft0, 0, out=cat16[..., 10:13])
        xt2 = t.warp(x1, ft1, 0, out=cat16[..., 13:16])
        mask_in = t.concat(cat16, [(ft0, 0, 2), (ft1, 2, 4), (xt1, 10, 13), (xt2, 13, 16)], consts=[(canvas, 4, 10)])
        mask = self._unet(t, mask_in, "Mask.", 4, final_act=ACT_SIGMOID)

        cat9 = t.concat_buffer(n, ch, cw, 9)                         # (input0, input1, output), :118
        out = t.blend(xt1, xt2, t.slice(mask, 0, 1), t.slice(mask, 1, 2), 1 - tt, tt, 1e-8, ops.BLEND_RATIO,
                      out=cat9[..., 6:9])
        fin = self._unet(t, t.concat(cat9, [(out, 6, 9)], consts=[(canvas, 0, 6)]), "final.", 4)
        res = t.clamp(t.add(fin, out), 0.0, 1.0)
        return t.to_nchw(res, top, left, height, width)

    # ------------------------------------------------------------------ reference plugin API
    def forward(self, input0, input1, t=0.5, params=None, **kwargs):
        self._t = float(t)
        try:
            return super().forward(input0, input1, params=params)
        finally:
            self._t = 0.5
