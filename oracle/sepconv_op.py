"""Oracle: adaptive separable convolution (test infrastructure, not product).

Restates the three CUDA-C kernels the reference launches on the hot path:

* ``kernel_Sepconv_updateOutput``         reference sepconv/sepconv_op/sepconv.py:5-30
* ``kernel_Sepconv_updateGradVertical``   reference sepconv/sepconv_op/sepconv.py:138-163
* ``kernel_Sepconv_updateGradHorizontal`` reference sepconv/sepconv_op/sepconv.py:165-190

``kernel_Sepconv_updateGradInput`` (:32-63) is never launched by the meta
system (the op's input is a data frame, ``needs_input_grad[0]`` is False) and
is not restated.

Two forms are given: ``*_loops`` is a literal pure-Python transcription of the
kernel loops (tiny cases only) and ``sepconv_forward`` / ``sepconv_backward``
are the vectorised forms used as the checker at real sizes.  The tests pin the
vectorised form against the literal one.
"""
import torch


def sepconv_forward_loops(inp, vertical, horizontal):
    """Literal transcription of kernel_Sepconv_updateOutput (sepconv.py:5-30)."""
    n, c, hi, wi = inp.shape
    f = min(vertical.shape[1], horizontal.shape[1])
    ho, wo = vertical.shape[2], vertical.shape[3]
    out = torch.zeros(n, c, ho, wo, dtype=inp.dtype)
    for s in range(n):
        for d in range(c):
            for y in range(ho):
                for x in range(wo):
                    acc = 0.0
                    for fy in range(f):
                        for fx in range(f):
                            acc += float(inp[s, d, y + fy, x + fx]) * float(vertical[s, fy, y, x]) \
                                   * float(horizontal[s, fx, y, x])
                    out[s, d, y, x] = acc
    return out


def sepconv_backward_loops(inp, vertical, horizontal, grad_out):
    """Literal transcription of updateGradVertical/Horizontal (sepconv.py:138-190)."""
    n, c, hi, wi = inp.shape
    f = vertical.shape[1]
    ho, wo = vertical.shape[2], vertical.shape[3]
    gv = torch.zeros_like(vertical)
    gh = torch.zeros_like(horizontal)
    for s in range(n):
        for k in range(f):
            for y in range(ho):
                for x in range(wo):
                    av = 0.0
                    ah = 0.0
                    for d in range(c):
                        go = float(grad_out[s, d, y, x])
                        for j in range(f):
                            av += go * float(inp[s, d, y + k, x + j]) * float(horizontal[s, j, y, x])
                            ah += go * float(inp[s, d, y + j, x + k]) * float(vertical[s, j, y, x])
                    gv[s, k, y, x] = av
                    gh[s, k, y, x] = ah
    return gv, gh


def sepconv_forward(inp, vertical, horizontal):
    """out[n,c,y,x] = sum_fy v[n,fy,y,x] * sum_fx in[n,c,y+fy,x+fx] * h[n,fx,y,x]
    (same sum as sepconv.py:20-26, factorised; SURVEY Appx E1)."""
    n, c, hi, wi = inp.shape
    f = vertical.shape[1]
    ho, wo = vertical.shape[2], vertical.shape[3]
    assert hi - f == ho - 1 and wi - f == wo - 1  # sepconv.py:266-267
    # windows along x: [n,c,hi,wo,f]
    win = inp.unfold(3, f, 1)
    out = torch.zeros(n, c, ho, wo, dtype=inp.dtype, device=inp.device)
    h = horizontal.permute(0, 2, 3, 1)  # [n,ho,wo,f]
    for fy in range(f):
        t = (win[:, :, fy:fy + ho] * h.unsqueeze(1)).sum(-1)  # [n,c,ho,wo]
        out = out + t * vertical[:, fy].unsqueeze(1)
    return out


def sepconv_backward(inp, vertical, horizontal, grad_out):
    """gV, gH of sepconv.py:138-190 in factorised form (SURVEY Appx E1)."""
    n, c, hi, wi = inp.shape
    f = vertical.shape[1]
    ho, wo = vertical.shape[2], vertical.shape[3]
    win = inp.unfold(3, f, 1)                      # [n,c,hi,wo,f]
    h = horizontal.permute(0, 2, 3, 1)             # [n,ho,wo,f]
    gv = torch.zeros_like(vertical)
    u = torch.zeros(n, c, ho, wo, f, dtype=inp.dtype, device=inp.device)
    for fy in range(f):
        w = win[:, :, fy:fy + ho]                  # [n,c,ho,wo,f]
        t = (w * h.unsqueeze(1)).sum(-1)           # [n,c,ho,wo]
        gv[:, fy] = (grad_out * t).sum(1)
        u = u + w * vertical[:, fy].unsqueeze(1).unsqueeze(-1)
    gh = (grad_out.unsqueeze(-1) * u).sum(1).permute(0, 3, 1, 2).contiguous()
    return gv, gh


class FunctionSepconvCPU(torch.autograd.Function):
    """CPU stand-in with the calling convention of the reference's
    ``FunctionSepconv`` (sepconv.py:247-380); gradInput is None as on the hot
    path (sepconv.py:319)."""

    @staticmethod
    def forward(ctx, inp, vertical, horizontal):
        ctx.save_for_backward(inp, vertical, horizontal)
        return sepconv_forward(inp, vertical, horizontal)

    @staticmethod
    def backward(ctx, grad_out):
        inp, vertical, horizontal = ctx.saved_tensors
        gv, gh = sepconv_backward(inp, vertical, horizontal, grad_out.contiguous())
        return None, gv, gh
