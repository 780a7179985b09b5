"""Minimal reverse-mode tape over the operator table.

The backbones are written once as forward code against ``Tape``; each op records
a closure that launches the matching backward kernels.  The tape is plain Python
that only enqueues kernels on the current stream, so a whole support step
(forward, loss, backward, fused update) can be captured into one CUDA graph and
replayed with no Python in the loop (decision 2).

Replaces torch.autograd on the hot path of the reference
(meta_learning_system.py:291-292 ``torch.autograd.grad`` over the backbone).
"""
from .ops import ACT_NONE


class Var:
    """An NHWC activation (or NCHW image for frames/predictions) plus its gradient slot."""
    __slots__ = ("data", "grad", "requires_grad", "act", "slope", "consumers", "grad_masked")

    def __init__(self, data, requires_grad=True):
        self.data = data
        self.grad = None
        self.requires_grad = requires_grad
        self.act = ACT_NONE        # activation fused into the producing conv's epilogue
        self.slope = 0.0
        self.consumers = 0         # ops reading this Var (decides whether the act mask can be fused)
        self.grad_masked = False   # True once .grad already is the gradient w.r.t. the PRE-activation


class ConvParam:
    """Weights of one conv as the kernels see them (KRSC weight view, bias vector)."""
    __slots__ = ("name", "w", "b", "_wt")

    def __init__(self, name, w, b):
        self.name, self.w, self.b = name, w, b
        self._wt = None

    def wt(self, ops):
        if self._wt is None:
            self._wt = ops.weight_to_dgrad(self.w)
        return self._wt


class Tape:
    def __init__(self, ops, params, sink=None):
        """params: callable name -> ConvParam; sink: object with weight_grad(param, x, dy, k) or None (no wgrad)."""
        self.ops = ops
        self.params = params
        self.sink = sink
        self.nodes = []

    # ------------------------------------------------------------------ helpers
    def _give(self, var, producer):
        """Hand a freshly computed gradient to ``var``: ``producer(out, accumulate)`` writes/accumulates it."""
        if not var.requires_grad:
            return
        if var.grad is None:
            var.grad = producer(None, False)
        else:
            producer(var.grad, True)

    def backward(self):
        for fn in reversed(self.nodes):
            fn()
        self.nodes = []

    # ------------------------------------------------------------------ ops
    def conv(self, x, name, act=ACT_NONE, slope=0.0, out=None):
        ops = self.ops
        p = self.params(name)
        k = p.w.shape[1]
        x.consumers += 1
        y = Var(ops.conv_fprop(x.data, p.w, p.b, act, slope, out=out))
        y.act, y.slope = act, slope

        def bwd():
            dy = y.grad
            if dy is None:
                return
            if act != ACT_NONE and not y.grad_masked:
                ops.act_bwd(dy, y.data, act, slope)
            # dgrad first: a fused-update sink may overwrite p.w in place, and the rotated copy
            # p.wt must come from the weights this forward pass actually used
            if x.requires_grad:
                if x.consumers == 1 and x.act != ACT_NONE and x.grad is None:
                    # sole consumer of an activated conv output: fold that activation's derivative into
                    # this dgrad's epilogue, so x.grad is born as the pre-activation gradient
                    x.grad = ops.conv_dgrad(dy, p.w, wt=p.wt(ops), mask_y=x.data, mask_act=x.act,
                                            mask_slope=x.slope)
                    x.grad_masked = True
                else:
                    self._give(x, lambda o, acc: ops.conv_dgrad(dy, p.w, wt=p.wt(ops), out=o, accumulate=acc))
            if self.sink is not None:
                self.sink.weight_grad(p, x.data, dy, k)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def avgpool(self, x):
        ops = self.ops
        x.consumers += 1
        y = Var(ops.avgpool_fwd(x.data))

        def bwd():
            if y.grad is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.empty_like_act(x.data)
                ops.avgpool_bwd(y.grad, x.grad, False)
            else:
                ops.avgpool_bwd(y.grad, x.grad, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def maxpool(self, x):
        ops = self.ops
        x.consumers += 1
        y = Var(ops.maxpool_fwd(x.data))

        def bwd():
            if y.grad is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.empty_like_act(x.data)
                ops.maxpool_bwd(x.data, y.grad, x.grad, False)
            else:
                ops.maxpool_bwd(x.data, y.grad, x.grad, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def upsample(self, x, align_corners, out=None):
        ops = self.ops
        x.consumers += 1
        y = Var(ops.upsample_fwd(x.data, align_corners, out=out))

        def bwd():
            if y.grad is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.empty_like_act(x.data)
                ops.upsample_bwd(y.grad, x.grad, align_corners, False)
            else:
                ops.upsample_bwd(y.grad, x.grad, align_corners, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def add(self, a, b):
        ops = self.ops
        a.consumers += 1
        b.consumers += 1
        y = Var(ops.add(a.data, b.data))

        def bwd():
            g = y.grad
            if g is None:
                return
            for v in (a, b):
                if not v.requires_grad:
                    continue
                if v.grad is None:
                    v.grad = ops.empty_like_act(v.data)
                    ops.copy(g, v.grad, False)
                else:
                    ops.copy(g, v.grad, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def concat_buffer(self, n, h, w, channels):
        """Allocate a concat target; producers write into channel slices (``out=`` of conv/upsample)."""
        return self.ops.empty_act(n, h, w, channels)

    def as_var_of_slices(self, buf, parts):
        """Var over a concat buffer whose channel slices were produced by ``parts`` (list of (Var, c0, c1))."""
        ops = self.ops
        y = Var(buf)
        for v, _, _ in parts:
            v.consumers += 1

        def bwd():
            g = y.grad
            if g is None:
                return
            for v, c0, c1 in parts:
                if not v.requires_grad:
                    continue
                gs = g[..., c0:c1]
                if v.grad is None:
                    v.grad = ops.empty_like_act(v.data)
                    ops.copy(gs, v.grad, False)
                else:
                    ops.copy(gs, v.grad, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def sepconv(self, frame, vert, horiz, oh, ow, gy0, gx0, iy0, ix0):
        """Adaptive separable convolution; ``frame`` is NCHW data, result is an NCHW Var."""
        ops = self.ops
        vert.consumers += 1
        horiz.consumers += 1
        y = Var(ops.sepconv_fwd(frame, vert.data, horiz.data, oh, ow, gy0, gx0, iy0, ix0))

        def bwd():
            g = y.grad
            if g is None:
                return
            n, gh, gw, taps = vert.data.shape
            assert vert.grad is None and horiz.grad is None, "sepconv filters have a single consumer"
            vert.grad = ops.zeros_act(n, gh, gw, taps)
            horiz.grad = ops.zeros_act(n, gh, gw, taps)
            ops.sepconv_bwd(frame, vert.data, horiz.data, g, vert.grad, horiz.grad, gy0, gx0, iy0, ix0)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def add_nchw(self, a, b):
        """a += b on contiguous NCHW outputs (sepconv/model.py:349 ``tensorDot1 + tensorDot2``); shares the gradient."""
        ops = self.ops
        a.consumers += 1
        b.consumers += 1
        ops.axpby(b.data, 1.0, a.data, 1.0)
        y = Var(a.data)

        def bwd():
            if y.grad is None:
                return
            a.grad = y.grad
            b.grad = y.grad

        self.nodes.append(bwd)
        return y
