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
    __slots__ = ("data", "grad", "requires_grad", "act", "slope", "consumers", "grad_masked", "clean", "grad_clean")

    def __init__(self, data, requires_grad=True):
        self.data = data
        self.grad = None
        self.requires_grad = requires_grad
        self.act = ACT_NONE        # activation fused into the producing conv's epilogue
        self.slope = 0.0
        self.consumers = 0         # ops reading this Var (decides whether the act mask can be fused)
        self.grad_masked = False   # True once .grad already is the gradient w.r.t. the PRE-activation
        self.clean = False         # True when .data is known to lie on the TF32 grid (its producer stored it rounded)
        self.grad_clean = False    # True when .grad was written once, rounded, by a kernel that takes the flag


class ConvParam:
    """Weights of one conv as the kernels see them: ``w`` the exact fp32 KRSC view (what updates read and write),
    ``wr`` the copy fprop reads and ``wt`` the rotated copy dgrad reads -- both rounded to the TF32 grid when the
    operator table follows the TF32 operand convention (include/mi_b200.h), else ``wr`` is ``w`` itself."""
    __slots__ = ("name", "w", "b", "_wt", "_wr")

    def __init__(self, name, w, b):
        self.name, self.w, self.b = name, w, b
        self._wt = None
        self._wr = None

    def wt(self, ops):
        if self._wt is None:
            self._wt = ops.weight_to_dgrad(self.w)
        return self._wt

    def wr(self, ops):
        if self._wr is None:
            if not ops.tf32_rn:
                self._wr = self.w
            else:
                cout, k, _, cin = self.w.shape
                self._wr = ops.round_tf32(self.w, out=ops.empty_weight(cout, cin, k))
        return self._wr


class Tape:
    def __init__(self, ops, params, sink=None, vectors=None):
        """params: callable name -> ConvParam; vectors: callable name -> (gamma, beta, mean, var, eps) of a frozen
        batch norm; sink: object with weight_grad(param, x, dy, k) [and bn_targets(name)] or None (no wgrad)."""
        self.ops = ops
        self.params = params
        self.vectors = vectors
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

    def _on_grid(self, x):
        """A conv is about to read ``x`` as a tensor-core operand: put it on the TF32 grid (in place, once) unless a
        conv epilogue already did.  Rounding is idempotent, so other readers of the buffer are unaffected beyond the
        rounding itself."""
        if self.ops.tf32_rn and not x.clean:
            self.ops.round_tf32(x.data)
            x.clean = True

    def _grad_on_grid(self, y, act, slope):
        """``y.grad`` is about to be the operand of this conv's dgrad / wgrad.  Born masked from the consumer conv's
        dgrad epilogue it already is on the grid; otherwise the (possibly trivial) activation-derivative pass rounds
        it on the way."""
        ops = self.ops
        if y.grad_masked or (act == ACT_NONE and y.grad_clean):
            return
        if act != ACT_NONE or ops.tf32_rn:
            ops.act_bwd(y.grad, y.data, act, slope, rnd=ops.tf32_rn)

    # ------------------------------------------------------------------ ops
    def conv(self, x, name, act=ACT_NONE, slope=0.0, out=None):
        ops = self.ops
        p = self.params(name)
        k = p.w.shape[1]
        x.consumers += 1
        self._on_grid(x)
        y = Var(ops.conv_fprop(x.data, p.wr(ops), p.b, act, slope, out=out))
        y.act, y.slope = act, slope
        y.clean = ops.tf32_rn

        def bwd():
            dy = y.grad
            if dy is None:
                return
            self._grad_on_grid(y, act, slope)
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
        y = Var(ops.avgpool_fwd(x.data, rnd=ops.tf32_rn))
        y.clean = ops.tf32_rn

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
        y.clean = x.clean          # a maximum of grid values is a grid value

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
        y = Var(ops.upsample_fwd(x.data, align_corners, out=out, rnd=ops.tf32_rn))
        y.clean = ops.tf32_rn

        def bwd():
            if y.grad is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.empty_like_act(x.data)
                if x.consumers == 1 and x.act != ACT_NONE:
                    # sole consumer of an activated conv output: its activation derivative is applied here, so
                    # x.grad is born as the pre-activation gradient and the conv skips its act_bwd pass
                    ops.upsample_bwd(y.grad, x.grad, align_corners, False, x.data, x.act, x.slope, rnd=ops.tf32_rn)
                    x.grad_masked = True       # (and on the TF32 grid: the producing conv's dgrad / wgrad read it next)
                else:
                    ops.upsample_bwd(y.grad, x.grad, align_corners, False)
            else:
                ops.upsample_bwd(y.grad, x.grad, align_corners, True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def crop(self, x, y0, x0, h, w):
        """Contiguous copy of the window x[:, y0:y0+h, x0:x0+w]; its gradient is zero outside the window."""
        ops = self.ops
        x.consumers += 1
        n, _, _, c = x.data.shape
        buf = ops.empty_act(n, h, w, c)
        ops.window_copy(x.data, (y0, x0), buf, (0, 0), (h, w))
        y = Var(buf, requires_grad=x.requires_grad)
        y.clean = x.clean

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.zeros_act(*x.data.shape)
            ops.window_copy(g, (0, 0), x.grad, (y0, x0), (h, w), accumulate=True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def upsample_window(self, x, align_corners, full_hw, lo_origin, hi_origin, hi_hw):
        """x2 bilinear upsample of a window of a larger grid, evaluated only on a window of the result."""
        ops = self.ops
        x.consumers += 1
        y = Var(ops.upsample_window_fwd(x.data, align_corners, full_hw, lo_origin, hi_origin, hi_hw, rnd=ops.tf32_rn))
        y.clean = ops.tf32_rn

        def bwd():
            if y.grad is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.empty_like_act(x.data)
                if x.consumers == 1 and x.act != ACT_NONE:
                    ops.upsample_window_bwd(y.grad, x.grad, align_corners, False, full_hw, lo_origin, hi_origin, x.data,
                                            x.act, x.slope, rnd=ops.tf32_rn)
                    x.grad_masked = True
                else:
                    ops.upsample_window_bwd(y.grad, x.grad, align_corners, False, full_hw, lo_origin, hi_origin)
            else:
                ops.upsample_window_bwd(y.grad, x.grad, align_corners, True, full_hw, lo_origin, hi_origin)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def add(self, a, b):
        ops = self.ops
        a.consumers += 1
        b.consumers += 1
        y = Var(ops.add(a.data, b.data, rnd=ops.tf32_rn))
        y.clean = ops.tf32_rn

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
        # on the TF32 grid when every slice was stored rounded and the slices tile the buffer (constants copied in by
        # `concat` are not tracked: their presence leaves the flag off)
        covered = sum(c1 - c0 for _, c0, c1 in parts) == buf.shape[-1]
        y.clean = covered and all(v.clean for v, _, _ in parts)

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
        n_img, taps_f = frame.shape[0], vert.data.shape[3]
        # tap-planar copies of the two filter tensors: written by the forward, read again by the backward
        planar = ops.sepconv_planar(n_img, oh, ow, taps_f) if (taps_f == 51 and frame.shape[1] == 3) else None
        y = Var(ops.sepconv_fwd(frame, vert.data, horiz.data, oh, ow, gy0, gx0, iy0, ix0, planar=planar))

        def bwd():
            g = y.grad
            if g is None:
                return
            n, gh, gw, taps = vert.data.shape
            assert vert.grad is None and horiz.grad is None, "sepconv filters have a single consumer"
            # fresh buffers: the op zero-fills outside the window in the launch that writes the window
            vert.grad = ops.empty_act(n, gh, gw, taps)
            horiz.grad = ops.empty_act(n, gh, gw, taps)
            scratch = ops.sepconv_planar(n_img, oh, ow, taps_f) if planar is not None else None
            ops.sepconv_bwd(frame, vert.data, horiz.data, g, vert.grad, horiz.grad, gy0, gx0, iy0, ix0,
                            rnd=ops.tf32_rn, planar=planar, planar_valid=planar is not None, planar_grad=scratch,
                            zero_outside=True)
            vert.grad_clean = horiz.grad_clean = ops.tf32_rn     # (zero outside the window is on the grid too)
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

    # ------------------------------------------------------------------ ops of the flow / attention backbones
    def _own_or_add(self, var, g):
        """Hand the freshly built gradient tensor ``g`` to ``var`` (ownership moves when it is the first one)."""
        if not var.requires_grad:
            return
        if var.grad is None:
            var.grad = g
        else:
            self.ops.copy(g, var.grad, True)

    def _grad_slot(self, var, zero=False):
        """(buffer, accumulate) to write ``var``'s gradient into; a fresh buffer is zeroed when ``zero``."""
        if var.grad is None:
            n, h, w, c = var.data.shape
            var.grad = self.ops.zeros_act(n, h, w, c) if zero else self.ops.empty_act(n, h, w, c)
            return var.grad, zero
        return var.grad, True

    def data(self, tensor):
        """A constant (frames, warped inputs): no gradient flows into it."""
        return Var(tensor, requires_grad=False)

    def slice(self, x, c0, c1):
        """Channel slice ``x[:, c0:c1]`` as a view (superslomo/model.py:597-598, rrin/model.py:83 ...)."""
        ops = self.ops
        x.consumers += 1
        y = Var(x.data[..., c0:c1], requires_grad=x.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            if x.grad is None:
                n, h, w, c = x.data.shape
                x.grad = ops.zeros_act(n, h, w, c)
            ops.copy(g, x.grad[..., c0:c1], True)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def lincomb(self, terms, const=0.0, out=None):
        """y = sum_i coef_i * x_i + const (flow interpolation coefficients, mask affine maps)."""
        ops = self.ops
        y_data = out
        for i, (coef, v) in enumerate(terms):
            v.consumers += 1
            y_data = ops.affine(v.data, coef, const if i == 0 else 0.0, out=y_data, accumulate=i > 0)
        y = Var(y_data, requires_grad=any(v.requires_grad for _, v in terms))

        def bwd():
            g = y.grad
            if g is None:
                return
            for coef, v in terms:
                if not v.requires_grad:
                    continue
                buf, acc = self._grad_slot(v)
                ops.affine(g, coef, 0.0, out=buf, accumulate=acc)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def warp(self, img, flow, variant, sx=1.0, sy=1.0, out=None):
        """Bilinear backward warp of the constant image ``img`` (NHWC data) by the Var ``flow``."""
        ops = self.ops
        flow.consumers += 1
        y = Var(ops.warp_fwd(img, flow.data, variant, sx, sy, out=out), requires_grad=flow.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not flow.requires_grad:
                return
            buf, acc = self._grad_slot(flow)
            ops.warp_bwd(img, flow.data, g, buf, variant, sx, sy, accumulate=acc)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def act(self, x, kind, slope=0.0):
        ops = self.ops
        x.consumers += 1
        y = Var(ops.act_fwd(x.data, kind, slope), requires_grad=x.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            ops.act_bwd(g, y.data, kind, slope)
            self._own_or_add(x, g)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def clamp(self, x, lo, hi):
        ops = self.ops
        x.consumers += 1
        y = Var(ops.clamp_fwd(x.data, lo, hi), requires_grad=x.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            buf, acc = self._grad_slot(x)
            ops.clamp_bwd(g, x.data, buf, lo, hi, acc)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def blend(self, a, b, m0, m1, w0, w1, eps, mode, out=None):
        """Visibility-weighted blend of two warped frames (include/mi_b200.h mi_blend_fwd)."""
        ops = self.ops
        parts = [v for v in (a, b, m0, m1) if v is not None]
        for v in parts:
            v.consumers += 1
        y = Var(ops.blend_fwd(a.data, b.data, m0.data, None if m1 is None else m1.data, w0, w1, eps, mode, out=out))

        def bwd():
            g = y.grad
            if g is None:
                return
            tg = []
            for v in (a, b, m0, m1):
                if v is None or not v.requires_grad:
                    tg.append(None)
                else:
                    tg.append(self._grad_slot(v, zero=True)[0])
            ops.blend_bwd(a.data, b.data, m0.data, None if m1 is None else m1.data, g, tg[0], tg[1], tg[2], tg[3],
                          True, w0, w1, eps, mode)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def to_nchw(self, x, y0, x0, h, w):
        """Crop window of an NHWC Var as the NCHW prediction (the paddingOutput of every backbone)."""
        ops = self.ops
        x.consumers += 1
        y = Var(ops.nhwc_window_to_nchw(x.data, y0, x0, h, w))

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            assert x.grad is None, "the cropped tensor has a single consumer"
            n, hh, ww, c = x.data.shape
            full = (y0 == 0 and x0 == 0 and hh == h and ww == w)
            x.grad = ops.empty_act(n, hh, ww, c) if full else ops.zeros_act(n, hh, ww, c)
            ops.nchw_to_nhwc_window(g.contiguous(), x.grad, y0, x0)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def to_nchw_shared(self, x, y0, x0, h, w):
        """``to_nchw`` for a Var that has other consumers too (the auxiliary outputs of SuperSloMo that feed the
        ``Super`` loss, superslomo/model.py:631-643): the cropped gradient is ADDED to whatever else reaches ``x``."""
        ops = self.ops
        x.consumers += 1
        y = Var(ops.nhwc_window_to_nchw(x.data, y0, x0, h, w), requires_grad=x.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            n, hh, ww, c = x.data.shape
            full = ops.zeros_act(n, hh, ww, c)
            ops.nchw_to_nhwc_window(g.contiguous(), full, y0, x0)
            self._own_or_add(x, full)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def from_nchw(self, x):
        """NCHW image Var -> NHWC activation Var (entry of a feature extractor applied to a prediction)."""
        ops = self.ops
        x.consumers += 1
        n, c, h, w = x.data.shape
        data = ops.empty_act(n, h, w, c)
        ops.nchw_to_nhwc_window(x.data.contiguous(), data, 0, 0)
        y = Var(data, requires_grad=x.requires_grad)

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            gn = ops.nhwc_window_to_nchw(g, 0, 0, h, w)
            if x.grad is None:
                x.grad = gn
            else:
                ops.axpby(gn, 1.0, x.grad, 1.0)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def concat(self, buf, parts, consts=()):
        """Var over the concat buffer ``buf``: ``parts`` = [(Var, c0, c1)] were produced in place (``out=`` slices),
        ``consts`` = [(tensor, c0, c1)] are constants copied in now."""
        for tsr, c0, c1 in consts:
            self.ops.copy(tsr, buf[..., c0:c1], False)
        return self.as_var_of_slices(buf, parts)

    def bn(self, x, name, act=ACT_NONE, slope=0.0, out=None):
        """Frozen batch norm + activation (voxel_flow.py:352-355: BN layers always run in eval mode)."""
        ops = self.ops
        gamma, beta, mean, var, eps = self.vectors(name)
        x.consumers += 1
        y = Var(ops.bn_eval_fwd(x.data, gamma, beta, mean, var, eps, act, slope, out=out))

        def bwd():
            dy = y.grad
            if dy is None:
                return
            tg = self.sink.bn_targets(name) if self.sink is not None else None
            dgamma, dbeta, mode, scale = tg if tg is not None else (None, None, 0, 1.0)
            dx, acc = (None, False)
            if x.requires_grad:
                dx, acc = self._grad_slot(x)
            if dx is not None or dgamma is not None:
                ops.bn_eval_bwd(dy, y.data, x.data, gamma, mean, var, eps, act, slope, dx, acc, dgamma, dbeta, mode,
                                scale)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def ring_conv(self, x, name, act=ACT_NONE, slope=0.0, ring_mode=1):
        """Convolution over a ringed buffer: the ring of ``x`` is filled in place (zeros or reflection), the
        zero-padding engine runs over the whole buffer and only the interior of the result is meaningful
        (MetaConvNorm = ReflectionPad2d(1) + conv, model_utils.py:821-849)."""
        ops = self.ops
        p = self.params(name)
        k = p.w.shape[1]
        x.consumers += 1
        ops.ring_fix(x.data, ring_mode)
        self._on_grid(x)
        y = Var(ops.conv_fprop(x.data, p.wr(ops), p.b, act, slope))
        y.act, y.slope = act, slope
        y.clean = ops.tf32_rn

        def bwd():
            dy = y.grad
            if dy is None:
                return
            ops.ring_fix(dy, ops.RING_ZERO)      # ring outputs are never consumed
            self._grad_on_grid(y, act, slope)
            if x.requires_grad:
                if x.consumers == 1 and x.act != ACT_NONE and x.grad is None:
                    g = ops.conv_dgrad(dy, p.w, wt=p.wt(ops), mask_y=x.data, mask_act=x.act, mask_slope=x.slope)
                    ops.ring_fold(g, ring_mode)
                    x.grad = g
                    x.grad_masked = True
                else:
                    g = ops.conv_dgrad(dy, p.w, wt=p.wt(ops))
                    ops.ring_fold(g, ring_mode)
                    self._own_or_add(x, g)
            if self.sink is not None:
                self.sink.weight_grad(p, x.data, dy, k)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def interior_mean(self, x, ring):
        """Global average pool over the interior of a ringed buffer -> [n,1,1,c] (MetaCALayer, model_utils.py:947)."""
        ops = self.ops
        n, h, w, c = x.data.shape
        scale = 1.0 / ((h - 2 * ring) * (w - 2 * ring))
        x.consumers += 1
        y = Var(ops.interior_reduce(x.data, None, ring, scale))

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            if x.grad is None:
                x.grad = ops.zeros_act(n, h, w, c)
            ops.interior_bcast_add(g, x.grad, ring, scale)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def scale_add(self, o, s, res, ring):
        """out = o * s[n,c] + res (channel attention rescale + residual, model_utils.py:955,985)."""
        ops = self.ops
        for v in (o, s, res):
            v.consumers += 1
        y = Var(ops.scale_add(o.data, s.data, res.data))

        def bwd():
            g = y.grad
            if g is None:
                return
            assert s.grad is None
            s.grad = ops.interior_reduce(g, o.data, ring, 1.0)
            if o.requires_grad:
                buf, acc = self._grad_slot(o)
                ops.scale_bwd(g, s.data, buf, acc)
            self._own_or_add(res, g)
            y.grad = None

        self.nodes.append(bwd)
        return y

    def depth_to_space(self, x, mean0, mean1, h, w, pad_top, pad_left, r):
        """Ringed NHWC features -> cropped NCHW image + mean shift (cain/model.py:84-94, model_utils.py:202-217)."""
        ops = self.ops
        x.consumers += 1
        y = Var(ops.depth_to_space(x.data, mean0, mean1, h, w, pad_top, pad_left, r))

        def bwd():
            g = y.grad
            if g is None or not x.requires_grad:
                return
            assert x.grad is None, "the shuffled tensor has a single consumer"
            x.grad = ops.empty_like_act(x.data)
            ops.depth_to_space_bwd(g.contiguous(), x.grad, pad_top, pad_left, r)
            y.grad = None

        self.nodes.append(bwd)
        return y
