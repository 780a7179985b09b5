"""Loss wrapper with the reference's interface (reference loss.py:278-350).

``Loss(args)`` parses ``'w*TYPE+...'`` and ``forward(sr, hr)`` returns
``{'<TYPE>': w*l, ..., 'total': sum}``.  L1 and MSE (the hot-path losses) are one
fused value+gradient kernel (``mi_loss_fwd_bwd``).  ``Super`` (reference loss.py:246-274,
the loss of scripts/run_superslomo.sh) is evaluated on the tape (``SuperTerms``; graph path:
fastpath.py, compat path: ``forward_with_plugin``): its reconstruction / warping / smoothness terms are ``mi_loss_fwd_bwd``
launches on crops of the SuperSloMo tape, its perceptual term runs VGG16 conv4_3 through
the same convolution engines on a side tape.  VGG / GAN / SSIM terms are outside SURVEY
section 8 and raise NotImplementedError.
"""
import torch
import torch.nn as nn

from .backbone import build_tape, collect_grads, default_ops
from .ops import ACT_NONE, ACT_RELU
from .tape import ConvParam, Tape, Var

# torchvision vgg16().features[:22] (loss.py:249-250): (features index, cin, cout), 'M' = 2x2 max-pool
VGG16_CONV4_3 = [(0, 3, 64), (2, 64, 64), 'M', (5, 64, 128), (7, 128, 128), 'M', (10, 128, 256), (12, 256, 256),
                 (14, 256, 256), 'M', (17, 256, 512), (19, 512, 512), (21, 512, 512)]


class SuperTerms:
    """The ``Super`` loss on a tape.  ``state`` holds torchvision's vgg16 ``features.N.weight/bias`` tensors."""
    RECN, WARP, PRCP = 204.0, 102.0, 0.005      # loss.py:272

    def __init__(self, ops, state):
        self.ops = ops
        self._params = {}
        for e in VGG16_CONV4_3:
            if e == 'M':
                continue
            idx, cin, cout = e
            w = ops.empty_weight(cout, cin, 3)
            w.copy_(state["features.%d.weight" % idx].to(ops.device).permute(0, 2, 3, 1))
            b = state["features.%d.bias" % idx].to(device=ops.device, dtype=torch.float32).contiguous()
            self._params["vgg.%d" % idx] = ConvParam("vgg.%d" % idx, w, b)

    def _features(self, tape, x):
        last = VGG16_CONV4_3[-1][0]
        for e in VGG16_CONV4_3:
            if e == 'M':
                x = tape.maxpool(x)
            else:
                x = tape.conv(x, "vgg.%d" % e[0], ACT_NONE if e[0] == last else ACT_RELU)
        return x

    def perceptual(self, pred, target, weight, loss_out, backward):
        """loss_out += weight * MSE(vgg(pred), vgg(target)); returns d/dpred (NCHW) when ``backward``."""
        ops = self.ops
        if pred.shape[2] % 8 or pred.shape[3] % 8:
            raise NotImplementedError("the Super loss needs frame sizes divisible by 8 (three 2x2 max-pools)")
        side = Tape(ops, self._params.__getitem__, sink=None)       # frozen weights: no weight gradients
        x = Var(pred, requires_grad=backward)
        feat = self._features(side, side.from_nchw(x))
        ref = self._features(side, side.from_nchw(Var(target, requires_grad=False)))
        grad = ops.empty_like_act(feat.data) if backward else None
        ops.loss_fwd_bwd(feat.data, ref.data, 1, weight, loss_out, grad)
        if not backward:
            side.nodes = []
            return None
        feat.grad = grad
        side.backward()
        return x.grad

    def _l1_pair(self, a, b, weight, loss_out, backward):
        """loss_out += weight * mean|a - b| for two contiguous tensors; returns d/da (d/db is its negative)."""
        ga = torch.empty_like(a) if backward else None
        self.ops.loss_fwd_bwd(a, b, 0, weight, loss_out, ga)
        return ga

    def _smooth(self, flow, weight, loss_out, backward):
        """weight * (mean|F[..., :-1] - F[..., 1:]| + mean|F[..., :-1, :] - F[..., 1:, :]|) (loss.py:268-270)."""
        g = torch.zeros_like(flow) if backward else None
        gx = self._l1_pair(flow[..., :-1].contiguous(), flow[..., 1:].contiguous(), weight, loss_out, backward)
        gy = self._l1_pair(flow[..., :-1, :].contiguous(), flow[..., 1:, :].contiguous(), weight, loss_out, backward)
        if backward:
            g[..., :-1] += gx
            g[..., 1:] -= gx
            g[..., :-1, :] += gy
            g[..., 1:, :] -= gy
        return g

    def seed(self, tape, out, aux, target, frame0, frame1, weight, loss_out, backward=True):
        """Add ``weight * Super`` (summed over the batch's samples: ``weight`` carries the sample count) to
        ``loss_out`` and seed the gradients of the prediction ``out`` and of the auxiliary tape Vars ``aux``."""
        ops = self.ops
        top, left, h, w = aux["window"]
        pred = out.data
        grad = torch.empty_like(pred) if backward else None
        ops.loss_fwd_bwd(pred, target, 0, weight * self.RECN, loss_out, grad)
        gp = self.perceptual(pred, target, weight * self.PRCP, loss_out, backward)
        if backward:
            ops.axpby(gp, 1.0, grad, 1.0)
            if out.grad is None:
                out.grad = grad
            else:                                   # pixel terms of the same loss string seeded it already
                ops.axpby(grad, 1.0, out.grad, 1.0)
        # warped intermediate frames against the target, warped input frames against the other input (:266)
        w0 = tape.warp(aux["i0"], aux["f10"], 0)
        w1 = tape.warp(aux["i1"], aux["f01"], 0)
        for var, ref in ((aux["g0"], target), (aux["g1"], target), (w0, frame1), (w1, frame0)):
            crop = tape.to_nchw_shared(var, top, left, h, w)
            g = self._l1_pair(crop.data, ref, weight * self.WARP, loss_out, backward)
            if backward:
                crop.grad = g
        for flow in (aux["f10"], aux["f01"]):
            crop = tape.to_nchw_shared(flow, top, left, h, w)
            g = self._smooth(crop.data, weight, loss_out, backward)
            if backward:
                crop.grad = g


class _PixelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, kind, ops):
        p = pred.detach().contiguous()
        t = target.detach().contiguous()
        out = torch.zeros(1, device=p.device, dtype=p.dtype)
        grad = torch.empty_like(p)
        ops.loss_fwd_bwd(p, t, kind, 1.0, out, grad)
        ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class _PluginLossFunction(torch.autograd.Function):
    """Compat-path form of a loss that differentiates through the plugin's auxiliary outputs (``Super``): the tape
    forward, ALL loss terms and their gradient seeds happen in ``forward``; ``backward`` runs the tape once and scales
    the parameter gradients by the incoming d/dtotal (first order, like every other path of this package)."""

    @staticmethod
    def forward(ctx, net, terms, specs, frame0, frame1, target, names, *tensors):
        ops = net.ops
        tape, out = build_tape(net, frame0, frame1, names, tensors)
        f0, f1, tgt = (t.detach().contiguous() for t in (frame0, frame1, target))
        values = []
        for kind, weight in specs:
            value = torch.zeros(1, device=out.data.device, dtype=torch.float32)
            if kind == 'Super':
                terms.seed(tape, out, net.aux_vars, tgt, f0, f1, weight, value, True)
            else:
                g = torch.empty_like(out.data)
                ops.loss_fwd_bwd(out.data, tgt, Loss.KINDS[kind], weight, value, g)
                if out.grad is None:
                    out.grad = g
                else:
                    ops.axpby(g, 1.0, out.grad, 1.0)
            values.append(value)
        ctx.tape, ctx.names, ctx.net = tape, names, net
        per_term = torch.cat(values)
        pred = out.data.clone() if net.clone_output else out.data
        ctx.mark_non_differentiable(pred, per_term)
        return per_term.sum(), pred, per_term

    @staticmethod
    def backward(ctx, g_total, g_pred, g_terms):
        grads = collect_grads(ctx.net, ctx.tape, ctx.names, ctx.needs_input_grad[7:])
        ctx.tape = None
        return (None,) * 7 + tuple(None if g is None else g * g_total for g in grads)


def load_vgg16_state(args):
    """torchvision's vgg16 weights for the Super loss: ``args.vgg16_weights`` (a saved vgg16 state_dict, for machines
    without network access) or torchvision's ImageNet weights exactly as loss.py:249 fetches them."""
    path = getattr(args, 'vgg16_weights', None)
    if isinstance(path, dict):
        return path
    if path is not None:
        return torch.load(path, map_location='cpu')
    from torchvision import models
    return models.vgg16(weights=models.VGG16_Weights.IMAGENET1K_V1).state_dict()


class Loss(nn.modules.loss._Loss):
    KINDS = {'L1': 0, 'MSE': 1}

    def __init__(self, args, ops=None):
        super().__init__()
        self.ops = ops
        self.loss = []
        self.super_terms = None
        for term in args.loss.split('+'):
            weight, loss_type = term.split('*')
            if loss_type == 'Super':
                if args.model != 'superslomo':
                    raise NotImplementedError('the Super loss needs the auxiliary outputs of the superslomo plugin')
                self.super_terms = SuperTerms(ops if ops is not None else default_ops(), load_vgg16_state(args))
            elif loss_type not in self.KINDS:
                raise NotImplementedError('loss %s is outside the B200 hot path (SURVEY section 8f)' % loss_type)
            self.loss.append({'type': loss_type, 'weight': float(weight)})

    def forward(self, sr, hr, **kwargs):
        ops = self.ops if self.ops is not None else default_ops()
        loss = 0
        losses = {}
        if self.super_terms is not None:
            raise NotImplementedError('the Super loss differentiates through the plugin\'s auxiliary outputs: use '
                                      'forward_with_plugin (net_forward does) instead of forward(sr, hr)')
        for l in self.loss:
            _loss = _PixelLoss.apply(sr, hr, self.KINDS[l['type']], ops)
            effective = l['weight'] * _loss
            losses[l['type']] = effective
            loss = loss + effective
        losses['total'] = loss
        return losses

    def forward_with_plugin(self, net, frame0, frame1, target, params=None):
        """Plugin forward + every loss term in one differentiable step (reference meta_learning_system.py:496-501 for
        superslomo: ``criterion(output[0], target, **output[1])``); returns (losses dict, prediction)."""
        names, tensors = net.resolve_params(params)
        specs = tuple((l['type'], l['weight']) for l in self.loss)
        total, pred, per_term = _PluginLossFunction.apply(net, self.super_terms, specs, frame0, frame1, target, names,
                                                          *tensors)
        losses = {l['type']: per_term[i] for i, l in enumerate(self.loss)}
        losses['total'] = total
        return losses, pred
