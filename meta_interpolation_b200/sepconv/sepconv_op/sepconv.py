"""Drop-in for the reference's ``sepconv/sepconv_op/sepconv.py:FunctionSepconv``.

Same call: ``FunctionSepconv.apply(input[N,C,Hi,Wi], vertical[N,F,Ho,Wo],
horizontal[N,F,Ho,Wo]) -> [N,C,Ho,Wo]`` with ``Hi - F == Ho - 1`` (sepconv.py:266-267),
contiguous NCHW fp32, no gradient for ``input`` (sepconv.py:319), ``NotImplementedError``
on CPU tensors (sepconv.py:293-294) -- but no cupy, no per-call source regex, and the
kernels are the smem-tiled sm_100a ones (csrc/sepconv.cu).
"""
import torch

from ...backbone import default_ops


class FunctionSepconv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, vertical, horizontal):
        if not input.is_cuda:
            raise NotImplementedError()
        ops = default_ops()
        n, c, hi, wi = input.shape
        f = min(vertical.size(1), horizontal.size(1))
        ho = min(vertical.size(2), horizontal.size(2))
        wo = min(vertical.size(3), horizontal.size(3))
        assert hi - f == ho - 1
        assert wi - f == wo - 1
        assert input.is_contiguous() and vertical.is_contiguous() and horizontal.is_contiguous()
        v = ops.empty_act(n, ho, wo, f)
        h = ops.empty_act(n, ho, wo, f)
        v.copy_(vertical.permute(0, 2, 3, 1))     # NCHW API boundary -> NHWC filter layout
        h.copy_(horizontal.permute(0, 2, 3, 1))
        ctx.save_for_backward(input, v, h)
        ctx.planar = ops.sepconv_planar(n, ho, wo, f) if (f == 51 and c == 3) else None
        return ops.sepconv_fwd(input, v, h, ho, wo, 0, 0, 0, 0, planar=ctx.planar)

    @staticmethod
    def backward(ctx, gradOutput):
        input, v, h = ctx.saved_tensors
        ops = default_ops()
        n, ho, wo, f = v.shape
        gv = ops.zeros_act(n, ho, wo, f)
        gh = ops.zeros_act(n, ho, wo, f)
        scratch = ops.sepconv_planar(n, ho, wo, f) if ctx.planar is not None else None
        ops.sepconv_bwd(input, v, h, gradOutput.contiguous(), gv, gh, 0, 0, 0, 0, planar=ctx.planar,
                        planar_valid=ctx.planar is not None, planar_grad=scratch)
        return None, gv.permute(0, 3, 1, 2).contiguous(), gh.permute(0, 3, 1, 2).contiguous()


class ModuleSepconv(torch.nn.Module):
    def forward(self, tensorFirst, tensorSecond, tensorThird):
        return FunctionSepconv.apply(tensorFirst, tensorSecond, tensorThird)
