"""Deep Voxel Flow plugin (drop-in for the reference's ``voxelflow/core/models/voxel_flow.py:MetaVoxelFlow``).

Same constructor, parameter names/shapes (23 tensors incl. the frozen batch-norm scale/shift),
``get_optim_policies()`` and ``forward(x0, x1, syn_type='inter', params=None, **kwargs)`` contract as reference
voxel_flow.py:227-530.  Compute: reflection padding to a multiple of 64 folded into the canvas kernel; conv (no
bias) -> frozen BN + ReLU in one pointwise kernel; encoder features written straight into the decoder's concat
buffers; the trilinear "voxel flow" sampling (:452-507) is two border-clamped align-corners bilinear gathers at
grid -/+ 0.5*flow plus one blend kernel with the mask 0.5*(1+m).
BN scale/shift are in the fast-weight dict but never read from it (:379,385...): they always use the
stored parameters and their gradients are skipped in support passes; BN always runs in eval mode (:352-355).
"""
import torch

from ....backbone import MetaBackbone
from ....ops import ACT_NONE, ACT_RELU, ACT_TANH
from ....padding import reflect_pads

_LAYERS = (("conv1", 6, 64, 5), ("conv2", 64, 128, 5), ("conv3", 128, 256, 3), ("bottleneck", 256, 256, 3),
           ("deconv1", 512, 256, 3), ("deconv2", 384, 128, 5), ("deconv3", 192, 64, 5))


class MetaVoxelFlow(MetaBackbone):
    def __init__(self, config=None, resume=False, ops=None):
        super().__init__(ops)
        self.config = config
        self.input_mean = [0.5 * 255, 0.5 * 255, 0.5 * 255]
        self.input_std = [0.5 * 255, 0.5 * 255, 0.5 * 255]
        self.syn_type = 'inter'
        # reference :241-274: each MetaConv2dLayer draws a xavier weight at construction; afterwards every conv
        # weight is redrawn N(0, 0.01) in module order (the RNG stream is reproduced draw for draw)
        for _, cin, cout, k in _LAYERS + (("conv4", 64, 3, 5),):
            torch.nn.init.xavier_uniform_(torch.empty(cout, cin, k, k))

        def init(name, shape):
            if len(shape) == 4:
                return torch.empty(*shape).normal_(0, 0.01)
            if name.endswith("_bn.weight"):
                return torch.ones(*shape)
            return torch.zeros(*shape)

        self._build_parameters(init)
        if resume:
            print('Loading model: pretrained_models/voxelflow_ft.pth')
            checkpoint = torch.load('pretrained_models/voxelflow_ft.pth')
            self.load_state_dict(checkpoint['state_dict'])
        if config is not None:
            setattr(self.config, 'mult_conv_w', [1, 1])
            setattr(self.config, 'mult_conv_b', [2, 0])
            setattr(self.config, 'mult_bn', [1, 1])

    # ------------------------------------------------------------------ structure
    def conv_specs(self):
        return [(name, cin, cout, k, False) for name, cin, cout, k in _LAYERS] + [("conv4", 64, 3, 5, True)]

    def param_entries(self):
        entries = []
        for name, cin, cout, k in _LAYERS:
            entries.append(("conv", name, cin, cout, k, False))
            entries.append(("bn", name + "_bn", cout))
        entries.append(("conv", "conv4", 64, 3, 5, True))
        return entries

    def is_routed(self, param_name):
        return "_bn." not in param_name

    def get_optim_policies(self):
        """reference :307-350: conv weights / conv biases / BN scale+shift parameter groups."""
        own = dict(self.named_parameters())
        weight = [own[n + ".weight"] for n in self.conv_names]
        bias = [own["conv4.bias"]]
        bn = [own[n + s] for n in self.bn_names for s in (".weight", ".bias")]
        return [{'params': weight, 'lr_mult': 1, 'decay_mult': 1, 'name': 'model weight'},
                {'params': bias, 'lr_mult': 2, 'decay_mult': 0, 'name': 'model bias'},
                {'params': bn, 'lr_mult': 1, 'decay_mult': 1, 'name': 'model bn scale/shift'}]

    # ------------------------------------------------------------------ graph
    def build_graph(self, t, frame0, frame1):
        """reference MetaVoxelFlow.forward :357-509; frames NCHW [n,3,H,W]; returns the NCHW prediction Var."""
        ops = t.ops
        n, _, height, width = frame0.shape
        left, right, top, bottom = reflect_pads(height, width, 6)
        ch, cw = height + top + bottom, width + left + right
        canvas = ops.frames_to_canvas(frame0, frame1, ch, cw, top, left, 1)
        i0, i1 = canvas[..., 0:3], canvas[..., 3:6]

        def block(x, name, out=None):
            return t.bn(t.conv(x, name, ACT_NONE), name + "_bn", ACT_RELU, out=out)

        cat1 = t.concat_buffer(n, ch // 4, cw // 4, 512)      # deconv1 input: (up(bottleneck), conv3)
        cat2 = t.concat_buffer(n, ch // 2, cw // 2, 384)      # deconv2 input: (up(deconv1), conv2)
        cat3 = t.concat_buffer(n, ch, cw, 192)                # deconv3 input: (up(deconv2), conv1)
        c1 = block(t.data(canvas), "conv1", out=cat3[..., 128:192])
        c2 = block(t.maxpool(c1), "conv2", out=cat2[..., 256:384])
        c3 = block(t.maxpool(c2), "conv3", out=cat1[..., 256:512])
        x = block(t.maxpool(c3), "bottleneck")
        u = t.upsample(x, False, out=cat1[..., 0:256])
        x = block(t.as_var_of_slices(cat1, [(u, 0, 256), (c3, 256, 512)]), "deconv1")
        u = t.upsample(x, False, out=cat2[..., 0:256])
        x = block(t.as_var_of_slices(cat2, [(u, 0, 256), (c2, 256, 384)]), "deconv2")
        u = t.upsample(x, False, out=cat3[..., 0:128])
        x = block(t.as_var_of_slices(cat3, [(u, 0, 128), (c1, 128, 192)]), "deconv3")
        x = t.conv(x, "conv4", ACT_TANH)

        flow, mask = t.slice(x, 0, 2), t.slice(x, 2, 3)
        o1 = t.warp(i0, flow, 1, -0.5, -0.5)                  # coor_1 = grid - 0.5*flow, :461-466
        o2 = t.warp(i1, flow, 1, 0.5, 0.5)                    # coor_2 = grid + 0.5*flow
        m = t.lincomb([(0.5, mask)], const=0.5)               # 0.5 * (1 + mask), :505
        out = t.blend(o1, o2, m, None, 1.0, 1.0, 0.0, ops.BLEND_LERP)
        return t.to_nchw(out, top, left, height, width)

    # ------------------------------------------------------------------ reference plugin API
    def forward(self, x0, x1, syn_type='inter', params=None, **kwargs):
        if syn_type != 'inter':
            raise ValueError('Unknown syn_type ' + syn_type)
        return super().forward(x0, x1, params=params)
