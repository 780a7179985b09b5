"""Loss wrapper with the reference's interface (reference loss.py:278-350).

``Loss(args)`` parses ``'w*TYPE+...'`` and ``forward(sr, hr)`` returns
``{'<TYPE>': w*l, ..., 'total': sum}``.  L1 and MSE (the hot-path losses) are one
fused value+gradient kernel (``mi_loss_fwd_bwd``).  VGG / GAN / SSIM / Super terms
are outside SURVEY section 8 for this round and raise NotImplementedError.
"""
import torch
import torch.nn as nn

from .backbone import default_ops


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


class Loss(nn.modules.loss._Loss):
    KINDS = {'L1': 0, 'MSE': 1}

    def __init__(self, args, ops=None):
        super().__init__()
        self.ops = ops
        self.loss = []
        for term in args.loss.split('+'):
            weight, loss_type = term.split('*')
            if loss_type not in self.KINDS:
                raise NotImplementedError('loss %s is outside the B200 hot path (SURVEY section 8f)' % loss_type)
            self.loss.append({'type': loss_type, 'weight': float(weight)})

    def forward(self, sr, hr, **kwargs):
        ops = self.ops if self.ops is not None else default_ops()
        loss = 0
        losses = {}
        for l in self.loss:
            _loss = _PixelLoss.apply(sr, hr, self.KINDS[l['type']], ops)
            effective = l['weight'] * _loss
            losses[l['type']] = effective
            loss = loss + effective
        losses['total'] = loss
        return losses
