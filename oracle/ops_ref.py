"""Oracle: per-operator CPU references mirroring ``meta_interpolation_b200.ops.CudaOps``
method for method (test infrastructure, not product).

Every method has the same name, argument meaning and layouts (NHWC activation
views, KRSC weight views) as the CUDA table, implemented with plain ATen CPU ops,
so that (a) ``tests/`` can compare each CUDA kernel with its reference on the
same inputs and (b) the host-side executor can be exercised on CPU by injecting
this table.  The product never imports this module.
"""
import math

import torch
import torch.nn.functional as F

from .sepconv_op import sepconv_forward, sepconv_backward

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
WG_STORE, WG_ACCUM, WG_SGD_SCALAR, WG_SGD_TENSOR = 0, 1, 2, 3


def pad4(c):
    return (c + 3) & ~3


def _nchw(t):
    return t.permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1)


def _oihw(w):
    return w.permute(0, 3, 1, 2)


def act_apply(v, act, slope):
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_LEAKY:
        return F.leaky_relu(v, slope)
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    if act == ACT_TANH:
        return torch.tanh(v)
    return v


def act_grad(y, act, slope):
    if act == ACT_RELU:
        return (y > 0).to(y.dtype)
    if act == ACT_LEAKY:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    if act == ACT_SIGMOID:
        return y * (1 - y)
    if act == ACT_TANH:
        return 1 - y * y
    return torch.ones_like(y)


class RefOps:
    name = "ref"

    def __init__(self, device="cpu", dtype=torch.float32, tf32_rn=False):
        self.device = torch.device(device)
        self.dtype = dtype
        self._launches = 0
        # exact fp32 arithmetic by default; tf32_rn=True only exercises the host logic of the TF32 operand convention
        # (rounded weight copies, rounding passes) -- the convolutions themselves stay exact
        self.tf32_rn = tf32_rn
        # producers that can store their result on the TF32 grid (rnd=True), like their CUDA counterparts
        for name, outs in (("avgpool_fwd", None), ("upsample_fwd", None), ("upsample_window_fwd", None),
                           ("add", None), ("frames_to_canvas", None), ("sepconv_bwd", (4, 5)),
                           ("upsample_bwd", (1,)), ("upsample_window_bwd", (1,))):
            setattr(self, name, self._rounding(getattr(self, name), outs))

    def _rounding(self, fn, outs):
        def wrapped(*a, rnd=False, **k):
            y = fn(*a, **k)
            if rnd:
                for t in ([y] if outs is None else [a[i] for i in outs]):
                    self.round_tf32(t)
            return y
        return wrapped

    def launch_count(self):
        return self._launches

    def set_workspace_slot(self, slot):
        pass

    # ------------------------------------------------------------------ allocation (same layouts as CudaOps)
    def empty_act(self, n, h, w, c, zero_pad=False):
        ld = pad4(c)
        buf = torch.zeros(n, h, w, ld, device=self.device, dtype=self.dtype)
        return buf[..., :c] if ld != c else buf

    zeros_act = empty_act

    def empty_like_act(self, t):
        n, h, w, c = t.shape
        return self.empty_act(n, h, w, c)

    def empty_weight(self, cout, cin, k):
        ld = pad4(cin)
        buf = torch.zeros(cout, k, k, ld, device=self.device, dtype=self.dtype)
        return buf[..., :cin] if ld != cin else buf

    # ------------------------------------------------------------------ convolution
    def conv_fprop(self, x, w, b, act=ACT_NONE, slope=0.0, out=None, engine=None):
        k = w.shape[1]
        y = F.conv2d(_nchw(x), _oihw(w), b, stride=1, padding=k // 2)
        y = _nhwc(act_apply(y, act, slope))
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    def begin_deferred_wgrad(self):
        pass              # (the CPU mirror finishes every weight gradient in place)

    def flush_deferred_wgrad(self):
        pass

    def set_sm_budget(self, ctas):
        return 0          # (a launch-geometry hint of the CUDA table; nothing to do on the CPU)

    def weight_to_dgrad(self, w, out=None, rnd=None):
        cout, k, _, cin = w.shape
        wt = out if out is not None else self.empty_weight(cin, cout, k)
        wt.copy_(torch.flip(w, dims=(1, 2)).permute(3, 1, 2, 0))
        if self.tf32_rn if rnd is None else rnd:
            self.round_tf32(wt)
        return wt

    def round_tf32(self, x, out=None):
        """cvt.rna.tf32.f32: round to nearest (ties away from zero) onto the 10-bit-mantissa grid."""
        y = x if out is None else out
        i = x.contiguous().view(torch.int32)
        r = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
        y.copy_(torch.where(torch.isfinite(x), r, x))
        return y

    def conv_dgrad(self, dy, w, wt=None, mask_y=None, mask_act=ACT_NONE, mask_slope=0.0, out=None, accumulate=False,
                   engine=None):
        k = w.shape[1]
        dx = _nhwc(F.conv_transpose2d(_nchw(dy), _oihw(w), None, stride=1, padding=k // 2))
        if mask_y is not None:
            dx = dx * act_grad(mask_y, mask_act, mask_slope)
        if out is None:
            out = self.empty_act(*dx.shape)
            accumulate = False
        if accumulate:
            out.add_(dx)
        else:
            out.copy_(dx)
        return out

    def conv_wgrad(self, x, dy, k, ldw, spec, engine=None):
        cin, cout = x.shape[3], dy.shape[3]
        gw = torch.nn.grad.conv2d_weight(_nchw(x), (cout, cin, k, k), _nchw(dy), stride=1, padding=k // 2)
        gw = gw.permute(0, 2, 3, 1)  # KRSC
        gb = dy.sum(dim=(0, 1, 2))
        has_b = (spec.grad_b is not None) if spec.mode <= WG_ACCUM else (spec.b_in is not None)
        if spec.mode == WG_STORE:
            spec.grad_w.copy_(gw)
            if has_b:
                spec.grad_b.copy_(gb)
        elif spec.mode == WG_ACCUM:
            spec.grad_w.add_(spec.scale * gw)
            if has_b:
                spec.grad_b.add_(spec.scale * gb)
        else:
            lr_w = spec.lr_w.reshape(-1)[0] if spec.mode == WG_SGD_SCALAR else spec.lr_w
            new_w = spec.w_in - lr_w * gw
            spec.w_out.copy_(new_w)
            rounded = getattr(spec, "wr_out", None) is not None
            if rounded:
                self.round_tf32(new_w.contiguous(), out=spec.wr_out)
            if getattr(spec, "wt_out", None) is not None:
                self.weight_to_dgrad(new_w, out=spec.wt_out, rnd=rounded)
            if spec.grad_w is not None:
                spec.grad_w.copy_(gw)
            if has_b:
                lr_b = spec.lr_b.reshape(-1)[0] if spec.mode == WG_SGD_SCALAR else spec.lr_b
                new_b = spec.b_in - lr_b * gb
                spec.b_out.copy_(new_b)
                if spec.grad_b is not None:
                    spec.grad_b.copy_(gb)
        if spec.gsum_w is not None:
            spec.gsum_w.add_(gw)
        if spec.gsum_b is not None:
            spec.gsum_b.add_(gb)

    # ------------------------------------------------------------------ resampling / pointwise
    def avgpool_fwd(self, x):
        y = _nhwc(F.avg_pool2d(_nchw(x), 2, 2))
        out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    def avgpool_bwd(self, dy, dx, accumulate):
        g = 0.25 * dy.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        dx.add_(g) if accumulate else dx.copy_(g)

    def maxpool_fwd(self, x):
        y = _nhwc(F.max_pool2d(_nchw(x), 2, 2))
        out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    @torch.enable_grad()
    def maxpool_bwd(self, x, dy, dx, accumulate):
        xx = _nchw(x).detach().clone().requires_grad_(True)
        y = F.max_pool2d(xx, 2, 2)
        (g,) = torch.autograd.grad(y, xx, _nchw(dy))
        g = _nhwc(g)
        dx.add_(g) if accumulate else dx.copy_(g)

    def upsample_fwd(self, x, align_corners, out=None):
        y = _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode="bilinear", align_corners=bool(align_corners)))
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    @torch.enable_grad()
    def upsample_bwd(self, dy, dx, align_corners, accumulate, mask_y=None, mask_act=ACT_NONE, mask_slope=0.0):
        n, h, w, c = dx.shape
        if mask_y is not None:
            return self.upsample_window_bwd(dy, dx, align_corners, accumulate, (h, w), (0, 0), (0, 0), mask_y,
                                            mask_act, mask_slope)
        xx = torch.zeros(n, c, h, w, dtype=dy.dtype, requires_grad=True)
        y = F.interpolate(xx, scale_factor=2, mode="bilinear", align_corners=bool(align_corners))
        (g,) = torch.autograd.grad(y, xx, _nchw(dy))
        g = _nhwc(g)
        dx.add_(g) if accumulate else dx.copy_(g)

    def upsample_window_fwd(self, x, align_corners, full_hw, lo_origin, hi_origin, hi_hw):
        """Reference = the FULL x2 upsampling of the window embedded in a zero grid, then cropped."""
        n, h, w, c = x.shape
        full = torch.zeros(n, full_hw[0], full_hw[1], c, dtype=x.dtype)
        full[:, lo_origin[0]:lo_origin[0] + h, lo_origin[1]:lo_origin[1] + w, :] = x
        up = _nhwc(F.interpolate(_nchw(full), scale_factor=2, mode="bilinear", align_corners=bool(align_corners)))
        y = self.empty_act(n, hi_hw[0], hi_hw[1], c)
        y.copy_(up[:, hi_origin[0]:hi_origin[0] + hi_hw[0], hi_origin[1]:hi_origin[1] + hi_hw[1], :])
        return y

    @torch.enable_grad()
    def upsample_window_bwd(self, dy, dx, align_corners, accumulate, full_hw, lo_origin, hi_origin, mask_y=None,
                            mask_act=ACT_NONE, mask_slope=0.0):
        n, h, w, c = dx.shape
        xin = torch.zeros(n, h, w, c, dtype=dx.dtype, requires_grad=True)
        full = F.pad(_nchw(xin), [lo_origin[1], full_hw[1] - lo_origin[1] - w, lo_origin[0],
                                  full_hw[0] - lo_origin[0] - h])
        up = F.interpolate(full, scale_factor=2, mode="bilinear", align_corners=bool(align_corners))
        win = up[:, :, hi_origin[0]:hi_origin[0] + dy.shape[1], hi_origin[1]:hi_origin[1] + dy.shape[2]]
        (g,) = torch.autograd.grad(win, xin, _nchw(dy))
        dx.add_(g) if accumulate else dx.copy_(g)
        if mask_y is not None:
            dx.mul_(act_grad(mask_y, mask_act, mask_slope))

    def window_copy(self, src, src_origin, dst, dst_origin, hw, accumulate=False):
        s = src[:, src_origin[0]:src_origin[0] + hw[0], src_origin[1]:src_origin[1] + hw[1], :]
        d = dst[:, dst_origin[0]:dst_origin[0] + hw[0], dst_origin[1]:dst_origin[1] + hw[1], :]
        d.add_(s) if accumulate else d.copy_(s)

    def add(self, a, b, out=None):
        if out is None:
            out = self.empty_act(*a.shape)
        out.copy_(a + b)
        return out

    def copy(self, src, dst, accumulate=False):
        dst.add_(src) if accumulate else dst.copy_(src)

    def act_bwd(self, dy, y, act, slope=0.0, rnd=False):
        dy.mul_(act_grad(y, act, slope))
        if rnd:
            self.round_tf32(dy)

    def fill(self, t, value):
        t.fill_(value)

    def axpby(self, x, a, y, b):
        y.copy_(a * x + (b * y if b != 0 else 0))

    # ------------------------------------------------------------------ glue ops of the flow-based backbones
    BIN_ADD, BIN_SUB, BIN_MUL, BIN_DIV = 0, 1, 2, 3

    def bn_eval_fwd(self, x, gamma, beta, mean, var, eps, act=ACT_NONE, slope=0.0, out=None):
        y = act_apply((x - mean) * torch.rsqrt(var + eps) * gamma + beta, act, slope)
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    def bn_eval_bwd(self, dy, y, x, gamma, mean, var, eps, act, slope, dx, accumulate_dx, dgamma, dbeta, mode, scale):
        dz = dy * act_grad(y, act, slope)
        inv = torch.rsqrt(var + eps)
        if dx is not None:
            g = dz * gamma * inv
            dx.add_(g) if accumulate_dx else dx.copy_(g)
        dg = (dz * (x - mean) * inv).sum(dim=(0, 1, 2))
        db = dz.sum(dim=(0, 1, 2))
        for tgt, val in ((dgamma, dg), (dbeta, db)):
            if tgt is not None:
                tgt.add_(scale * val) if mode == WG_ACCUM else tgt.copy_(val)

    @staticmethod
    def _bin(op, a, b):
        return (a + b, a - b, a * b, a / b)[op]

    def binary_fwd(self, op, a, b, out=None):
        y = self._bin(op, a, b)
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    def binary_bwd(self, op, a, b, go, ga, acc_a, gb, acc_b):
        if op == 0:
            da, db = go, go
        elif op == 1:
            da, db = go, -go
        elif op == 2:
            da, db = go * b, go * a
        else:
            da, db = go / b, -go * a / (b * b)
        if b.shape[3] == 1 and a.shape[3] != 1:
            db = db.sum(dim=3, keepdim=True)
        if ga is not None:
            ga.add_(da) if acc_a else ga.copy_(da)
        if gb is not None:
            gb.add_(db) if acc_b else gb.copy_(db)

    def affine(self, x, alpha, beta, out=None, accumulate=False):
        y = alpha * x + beta
        if out is None:
            out = self.empty_act(*y.shape)
            accumulate = False
        out.add_(y) if accumulate else out.copy_(y)
        return out

    # ------------------------------------------------------------------ heads of the flow / attention backbones
    BLEND_RATIO, BLEND_RATIO_COMPLEMENT, BLEND_LERP = 0, 1, 2
    RING_ZERO, RING_REFLECT = 0, 1

    def _out(self, y, out):
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    def act_fwd(self, x, act, slope=0.0, out=None):
        return self._out(act_apply(x, act, slope), out)

    def clamp_fwd(self, x, lo, hi, out=None):
        return self._out(x.clamp(lo, hi), out)

    def clamp_bwd(self, dy, x, dx, lo, hi, accumulate):
        g = dy * ((x >= lo) & (x <= hi)).to(dy.dtype)
        dx.add_(g) if accumulate else dx.copy_(g)

    @staticmethod
    def _blend(a, b, m0, m1, w0, w1, eps, mode):
        if mode == 2:
            return m0 * a + (1 - m0) * b
        if mode == 1:
            m1 = 1 - m0
        return (w0 * m0 * a + w1 * m1 * b) / (w0 * m0 + w1 * m1 + eps)

    def blend_fwd(self, a, b, m0, m1, w0, w1, eps, mode, out=None):
        return self._out(self._blend(a, b, m0, m1, w0, w1, eps, mode), out)

    @torch.enable_grad()
    def blend_bwd(self, a, b, m0, m1, go, ga, gb, gm0, gm1, accumulate, w0, w1, eps, mode):
        ins = [t.detach().clone().requires_grad_(True) for t in (a, b, m0)]
        m1v = m1.detach().clone().requires_grad_(True) if (mode == 0) else None
        y = self._blend(ins[0], ins[1], ins[2], m1v, w0, w1, eps, mode)
        wrt = ins + ([m1v] if m1v is not None else [])
        grads = torch.autograd.grad(y, wrt, go)
        for tgt, g in zip((ga, gb, gm0, gm1), list(grads) + [None] * (4 - len(grads))):
            if tgt is not None and g is not None:
                tgt.add_(g) if accumulate else tgt.copy_(g)

    def ring_fix(self, x, mode):
        inner = _nchw(x[:, 1:-1, 1:-1, :])
        y = F.pad(inner, [1, 1, 1, 1], mode="reflect") if mode == 1 else F.pad(inner, [1, 1, 1, 1])
        x.copy_(_nhwc(y))

    @torch.enable_grad()
    def ring_fold(self, g, mode):
        n, h, w, c = g.shape
        if mode == 1:
            inner = torch.zeros(n, c, h - 2, w - 2, dtype=g.dtype, requires_grad=True)
            y = F.pad(inner, [1, 1, 1, 1], mode="reflect")
            (gi,) = torch.autograd.grad(y, inner, _nchw(g))
        else:
            gi = _nchw(g[:, 1:-1, 1:-1, :])
        g.copy_(_nhwc(F.pad(gi.detach(), [1, 1, 1, 1])))

    def channel_mean_nchw(self, f):
        return f.mean(2).mean(2).reshape(-1)

    def space_to_depth(self, f0, f1, mean0, mean1, pad_top, pad_left, oh, ow, r):
        n, c, h, w = f0.shape
        pads = [pad_left, ow * r - w - pad_left, pad_top, oh * r - h - pad_top]
        outs = []
        for f, m in ((f0, mean0), (f1, mean1)):
            x = f - m.view(n, c, 1, 1)
            if any(pads):
                x = F.pad(x, pads, mode="reflect")
            v = x.contiguous().view(n, c, oh, r, ow, r).permute(0, 1, 3, 5, 2, 4).contiguous()
            outs.append(v.view(n, c * r * r, oh, ow))
        y = F.pad(torch.cat(outs, 1), [1, 1, 1, 1])
        out = self.empty_act(n, oh + 2, ow + 2, 2 * c * r * r)
        out.copy_(_nhwc(y))
        return out

    def depth_to_space(self, x, mean0, mean1, h, w, pad_top, pad_left, r):
        n, hh, ww, c = x.shape
        ih, iw = hh - 2, ww - 2
        inner = _nchw(x[:, 1:-1, 1:-1, :]).contiguous()
        oc = c // (r * r)
        y = inner.view(n, oc, r, r, ih, iw).permute(0, 1, 4, 2, 5, 3).contiguous().view(n, oc, ih * r, iw * r)
        y = y[:, :, pad_top:pad_top + h, pad_left:pad_left + w]
        return (y + 0.5 * (mean0 + mean1).view(n, oc, 1, 1)).contiguous()

    def depth_to_space_bwd(self, gout, gin, pad_top, pad_left, r):
        n, oc, h, w = gout.shape
        _, hh, ww, c = gin.shape
        ih, iw = hh - 2, ww - 2
        full = torch.zeros(n, oc, ih * r, iw * r, dtype=gout.dtype)
        full[:, :, pad_top:pad_top + h, pad_left:pad_left + w] = gout
        v = full.view(n, oc, ih, r, iw, r).permute(0, 1, 3, 5, 2, 4).contiguous().view(n, c, ih, iw)
        gin.copy_(_nhwc(F.pad(v, [1, 1, 1, 1])))

    def interior_reduce(self, x, mul, ring, scale):
        n, h, w, c = x.shape
        v = x if mul is None else x * mul
        if ring:
            v = v[:, ring:h - ring, ring:w - ring, :]
        out = self.empty_act(n, 1, 1, c)
        out.copy_(scale * v.sum(dim=(1, 2), keepdim=True))
        return out

    def scale_add(self, o, s, res, out=None):
        y = o * s
        if res is not None:
            y = y + res
        return self._out(y, out)

    def scale_bwd(self, g, s, dx, accumulate):
        v = g * s
        dx.add_(v) if accumulate else dx.copy_(v)

    def interior_bcast_add(self, dy, dx, ring, scale):
        n, h, w, c = dx.shape
        dx[:, ring:h - ring, ring:w - ring, :].add_(dy * scale)

    # ------------------------------------------------------------------ frames in / prediction out
    def frames_to_canvas(self, f0, f1, ch, cw, pad_top, pad_left, mode):
        n, c, h, w = f0.shape
        pad = [pad_left, cw - pad_left - w, pad_top, ch - pad_top - h]
        m = "replicate" if mode == 0 else "reflect"
        x = torch.cat([F.pad(f0, pad, mode=m), F.pad(f1, pad, mode=m)], 1)
        out = self.empty_act(n, ch, cw, 6)
        out.copy_(_nhwc(x))
        return out

    def nhwc_window_to_nchw(self, src, y0, x0, h, w):
        return _nchw(src[:, y0:y0 + h, x0:x0 + w, :]).contiguous()

    def nchw_to_nhwc_window(self, src, dst, y0, x0):
        n, c, h, w = src.shape
        dst[:, y0:y0 + h, x0:x0 + w, :].copy_(_nhwc(src))

    # ------------------------------------------------------------------ adaptive separable convolution
    @staticmethod
    def _sep_input(frame, oh, ow, iy0, ix0, taps):
        n, c, fh, fw = frame.shape
        ys = (torch.arange(oh + taps - 1) + iy0).clamp(0, fh - 1)
        xs = (torch.arange(ow + taps - 1) + ix0).clamp(0, fw - 1)
        return frame[:, :, ys][:, :, :, xs]

    def sepconv_planar(self, n, oh, ow, taps):
        return None           # (a workspace of the CUDA kernels; the restatement needs none)

    def sepconv_fwd(self, frame, vert, horiz, oh, ow, gy0, gx0, iy0, ix0, planar=None):
        taps = vert.shape[3]
        inp = self._sep_input(frame, oh, ow, iy0, ix0, taps)
        v = _nchw(vert[:, gy0:gy0 + oh, gx0:gx0 + ow, :])
        h = _nchw(horiz[:, gy0:gy0 + oh, gx0:gx0 + ow, :])
        return sepconv_forward(inp, v, h)

    def sepconv_bwd(self, frame, vert, horiz, grad_out, g_vert, g_horiz, gy0, gx0, iy0, ix0, planar=None,
                    planar_valid=False, planar_grad=None, zero_outside=False):
        if zero_outside:   # fresh buffers: the op defines the whole grids (MI_SEPCONV_ZERO_OUTSIDE of the C ABI)
            g_vert.zero_()
            g_horiz.zero_()
        taps = vert.shape[3]
        oh, ow = grad_out.shape[2], grad_out.shape[3]
        inp = self._sep_input(frame, oh, ow, iy0, ix0, taps)
        v = _nchw(vert[:, gy0:gy0 + oh, gx0:gx0 + ow, :]).contiguous()
        h = _nchw(horiz[:, gy0:gy0 + oh, gx0:gx0 + ow, :]).contiguous()
        gv, gh = sepconv_backward(inp, v, h, grad_out)
        g_vert[:, gy0:gy0 + oh, gx0:gx0 + ow, :].copy_(_nhwc(gv))
        g_horiz[:, gy0:gy0 + oh, gx0:gx0 + ow, :].copy_(_nhwc(gh))

    # ------------------------------------------------------------------ warp
    @staticmethod
    def _warp_grid(flow, variant, sx, sy):
        n, h, w, _ = flow.shape
        u, v = flow[..., 0], flow[..., 1]
        if variant == 0:   # superslomo/model.py:292-302, rrin/model.py:8-21
            gx = torch.arange(w, dtype=flow.dtype).view(1, 1, w).expand(n, h, w)
            gy = torch.arange(h, dtype=flow.dtype).view(1, h, 1).expand(n, h, w)
            x = 2 * ((gx + sx * u) / w - 0.5)
            y = 2 * ((gy + sy * v) / h - 0.5)
            return torch.stack((x, y), dim=3), dict(mode="bilinear", padding_mode="zeros", align_corners=False)
        gx = torch.linspace(-1.0, 1.0, w).view(1, 1, w).expand(n, h, w)
        gy = torch.linspace(-1.0, 1.0, h).view(1, h, 1).expand(n, h, w)
        return torch.stack((gx + sx * u, gy + sy * v), dim=3), dict(mode="bilinear", padding_mode="border",
                                                                    align_corners=True)

    def warp_fwd(self, img, flow, variant, sx=1.0, sy=1.0, out=None):
        grid, kw = self._warp_grid(flow, variant, sx, sy)
        y = _nhwc(F.grid_sample(_nchw(img), grid, **kw))
        if out is None:
            out = self.empty_act(*y.shape)
        out.copy_(y)
        return out

    @torch.enable_grad()
    def warp_bwd(self, img, flow, grad_out, grad_flow, variant, sx=1.0, sy=1.0, accumulate=False):
        fl = flow.detach().clone().contiguous().requires_grad_(True)
        grid, kw = self._warp_grid(fl, variant, sx, sy)
        y = F.grid_sample(_nchw(img), grid, **kw)
        (g,) = torch.autograd.grad(y, fl, _nchw(grad_out))
        grad_flow.add_(g) if accumulate else grad_flow.copy_(g)

    # ------------------------------------------------------------------ loss / metrics / optimizers
    def loss_fwd_bwd(self, pred, target, kind, weight, loss_out, grad=None):
        d = pred - target
        cnt = pred.numel()
        if kind == 0:
            loss_out.add_(weight * d.abs().mean())
            if grad is not None:
                grad.copy_(weight * torch.sign(d) / cnt)
        else:
            loss_out.add_(weight * (d * d).mean())
            if grad is not None:
                grad.copy_(weight * 2 * d / cnt)

    def psnr_accumulate(self, pred, target, sq_out):
        q = lambda t: t.mul(255).clamp(0, 255).round()
        d = (q(pred) - q(target)).div(255)
        sq_out.add_(d.pow(2).double().sum())

    def ssim_accumulate(self, pred, target, window, sum_out, val_range=255.0):
        """pytorch_msssim/__init__.py:19-75 on the 8-bit quantised images (utils.py:195-204); sum of the SSIM map."""
        q = lambda t: t.mul(255).clamp(0, 255).round()
        a, b = q(pred).unsqueeze(0), q(target).unsqueeze(0)
        c = a.shape[1]
        w2 = window.unsqueeze(1).mm(window.unsqueeze(0)).unsqueeze(0).unsqueeze(0).expand(c, 1, -1, -1).contiguous()
        F = torch.nn.functional
        mu1, mu2 = F.conv2d(a, w2, groups=c), F.conv2d(b, w2, groups=c)
        s1 = F.conv2d(a * a, w2, groups=c) - mu1.pow(2)
        s2 = F.conv2d(b * b, w2, groups=c) - mu2.pow(2)
        s12 = F.conv2d(a * b, w2, groups=c) - mu1 * mu2
        c1, c2 = (0.01 * val_range) ** 2, (0.03 * val_range) ** 2
        v1, v2 = 2.0 * s12 + c2, s1 + s2 + c2
        sum_out.add_((((2 * mu1 * mu2 + c1) * v1) / ((mu1.pow(2) + mu2.pow(2) + c1) * v2)).double().sum())

    def septuplet_prepare(self, src, y0, x0, reversed_, h, w, bgr=True, div255=True, mean=None, std=None):
        """data/vimeo_septuplet.py:50-78 frame by frame: crop (:59-61), temporal flip (:64-66), BGR->RGB (:69),
        HWC->CHW float (/255 unless voxelflow, :72-75), Normalize = (x - mean) / std (:77-78)."""
        tasks, frames = src.shape[:2]
        out = torch.empty(frames, tasks, 3, h, w, dtype=torch.float32)
        for b in range(tasks):
            order = list(range(frames))[::-1] if int(reversed_[b]) else list(range(frames))
            for f, fs in enumerate(order):
                im = src[b, fs, int(y0[b]):int(y0[b]) + h, int(x0[b]):int(x0[b]) + w, :]
                if bgr:
                    im = im[:, :, [2, 1, 0]]
                t = im.permute(2, 0, 1).contiguous().float()
                if div255:
                    t = t / 255
                if mean is not None:
                    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
                    s = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
                    t = (t - m) / s
                out[f, b] = t
        return out

    def inner_update(self, w_in, g, w_out, exp_avg, exp_avg_sq, lr, lr_per_element, lr_stride, num_step, seg, skip,
                     rule, step_count):
        n = w_in.numel()
        t = seg.long().repeat_interleave(1024)[:n]
        valid = t >= 0
        if skip is not None:
            valid = valid & ~(skip.bool()[t.clamp(min=0)])
        l = lr.reshape(-1) if lr_per_element else lr.reshape(-1)[(t.clamp(min=0) * lr_stride + num_step)]
        b1, b2, eps = 0.9, 0.99, 1e-8
        bc1 = 1 - b1 ** step_count
        bc2 = 1 - b2 ** step_count
        wi, gi = w_in.reshape(-1), g.reshape(-1)
        if rule == 0:
            o = wi - l * gi
        elif rule == 1:
            m = b1 * exp_avg.reshape(-1) + (1 - b1) * gi
            v = b2 * exp_avg_sq.reshape(-1) + (1 - b2) * gi * gi
            exp_avg.reshape(-1)[valid] = m[valid]
            exp_avg_sq.reshape(-1)[valid] = v[valid]
            o = wi - (l / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
        elif rule == 2:
            m = b1 * exp_avg.reshape(-1) + (1 - b1) * gi
            exp_avg.reshape(-1)[valid] = m[valid]
            o = wi - (l / bc1) * m / (gi.abs() + eps)
        else:
            o = wi - (l / bc1) * ((1 - b1) * gi) / (gi.abs() + eps)
        w_out.reshape(-1).copy_(torch.where(valid, o, wi))

    def outer_step(self, p, g, m, v, kind, lr, beta1, beta2, eps, weight_decay, step):
        gi = g + weight_decay * p if weight_decay != 0 else g
        if kind == 0:
            p.sub_(lr * gi)
        elif kind == 1:
            m.mul_(beta1).add_(gi, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
            denom = (v.sqrt() / math.sqrt(1 - beta2 ** step)).add_(eps)
            p.addcdiv_(m, denom, value=-(lr / (1 - beta1 ** step)))
        else:
            m.mul_(beta1).add_(gi, alpha=1 - beta1)
            torch.maximum(v * beta2, gi.abs() + eps, out=v)
            p.addcdiv_(m, v, value=-(lr / (1 - beta1 ** step)))

    def addcmul(self, y, a, x1, x2):
        y.add_(a * x1 * x2)

    def segment_dot(self, a, b, seg, out):
        n = a.numel()
        t = seg.long().repeat_interleave(1024)[:n]
        prod = a.reshape(-1) if b is None else (a.reshape(-1) * b.reshape(-1))
        valid = t >= 0
        out.index_add_(0, t[valid], prod[valid])

    def segment_scale(self, x, scale, seg, mask, y, alpha=1.0, accumulate=False):
        n = x.numel()
        t = seg.long().repeat_interleave(1024)[:n]
        valid = t >= 0
        tc = t.clamp(min=0)
        sc = scale.reshape(-1)[tc]
        if mask is not None:
            sc = torch.where(mask.reshape(-1)[tc] != 0, sc, torch.ones_like(sc))
        v = alpha * sc * x.reshape(-1)
        yy = y.reshape(-1)
        yy[valid] = (yy[valid] + v[valid]) if accumulate else v[valid]
