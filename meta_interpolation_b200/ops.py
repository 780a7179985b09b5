"""Operator table of the hot path: thin tensor-level wrappers over the C ABI.

Tensors are torch CUDA tensors used purely as device-memory handles:

* activation: view of shape [N,H,W,C] with strides (H*W*ld, W*ld, ld, 1); ``ld``
  (= ``t.stride(2)``) may exceed C (channel padding to 4, or a channel slice of a
  concat buffer);
* conv weight: "KRSC" view of shape [Cout,k,k,Cin] with ``stride(2) == ldw``;
  the reference's OIHW parameter is ``w.permute(0,3,1,2)`` of the same storage.

``CudaOps`` is the only operator table the product ships.  Host-side logic
(tape, backbones, meta system) is written against this interface so the CPU
test-suite can inject the oracle's table (oracle/ops_ref.py) to check the host
logic without a GPU; nothing in this package imports the oracle.
"""
import os
import weakref

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
_POISON = os.environ.get("MI_B200_POISON", "") == "1"   # fill fresh activation / workspace buffers with NaN
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
WG_STORE, WG_ACCUM, WG_SGD_SCALAR, WG_SGD_TENSOR = 0, 1, 2, 3


def pad4(c):
    return (c + 3) & ~3


def _ld(t):
    """pixel stride of an NHWC activation view (floats)."""
    assert t.dim() == 4 and t.stride(3) == 1, "NHWC view with unit channel stride expected"
    n, h, w, _ = t.shape
    ld = t.stride(2)
    assert h == 1 or t.stride(1) == w * ld, "rows must be contiguous in NHWC"
    assert n == 1 or t.stride(0) == h * w * ld, "images must be contiguous in NHWC"
    return ld


def _ldw(w):
    assert w.dim() == 4 and w.stride(3) == 1
    co, k, k2, ci = w.shape
    ld = w.stride(2)
    assert k == k2 and (k == 1 or w.stride(1) == k * ld) and w.stride(0) == k * k * ld, "KRSC weight view expected"
    return ld


class WgradSpec:
    """What the weight-gradient finishing stage does (include/mi_b200.h MI_WG_*)."""
    __slots__ = ("mode", "scale", "grad_w", "grad_b", "w_in", "b_in", "w_out", "b_out", "lr_w", "lr_b", "gsum_w",
                 "gsum_b", "wt_out", "wr_out")

    def __init__(self, mode=WG_STORE, scale=1.0, grad_w=None, grad_b=None, w_in=None, b_in=None, w_out=None,
                 b_out=None, lr_w=None, lr_b=None, gsum_w=None, gsum_b=None, wt_out=None, wr_out=None):
        self.wt_out = wt_out        # SGD modes: also emit the updated weight in the dgrad (rotated) layout
        self.wr_out = wr_out        # SGD modes: TF32-rounded copy of the updated weight (what the next fprop reads)
        self.mode, self.scale = mode, scale
        self.grad_w, self.grad_b = grad_w, grad_b
        self.w_in, self.b_in, self.w_out, self.b_out = w_in, b_in, w_out, b_out
        self.lr_w, self.lr_b, self.gsum_w, self.gsum_b = lr_w, lr_b, gsum_w, gsum_b


class CudaOps:
    """sm_100a operator table (libmi_b200.so).  Raises if the library is absent."""

    name = "cuda"

    def __init__(self, device=None, engine=ENGINE_AUTO):
        self.lib = _lib.load()
        self._padded = {}        # data_ptr of a padded activation buffer -> (row stride, channels)
        if not torch.cuda.is_available():
            raise _lib.MiB200Error("CudaOps needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        env = os.environ.get("MI_B200_ENGINE", "").lower()   # debugging aid: force one conv engine everywhere
        self.engine = {"simt": ENGINE_SIMT, "tc": ENGINE_TC}.get(env, engine)
        # TF32 operand convention (include/mi_b200.h): with a tensor-core engine every conv operand is kept on the TF32
        # grid by round-to-nearest at its producer; the exact-fp32 engine rounds nothing
        self.tf32_rn = self.engine != ENGINE_SIMT and os.environ.get("MI_B200_TF32_RN", "1") != "0"
        self._ws = None
        self.replayed_launches = 0   # kernels executed through CUDA-graph replays (counted at capture time)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty_act(self, n, h, w, c, zero_pad=False):
        ld = pad4(c)
        buf = torch.empty(n, h, w, ld, device=self.device, dtype=torch.float32)
        if _POISON:      # debugging aid: a kernel that reads what nobody wrote turns the result into NaN
            buf.fill_(float("nan"))
        if zero_pad and ld != c:
            buf[..., c:].zero_()
        return self._padded_view(buf, c)

    def zeros_act(self, n, h, w, c):
        buf = torch.zeros(n, h, w, pad4(c), device=self.device, dtype=torch.float32)
        return self._padded_view(buf, c)

    def _padded_view(self, buf, c):
        """[..., :c] of a buffer this table allocated with 1-3 pad lanes per row.  The allocation is remembered (until
        the buffer dies), so a tensor-core convolution writing into the view may be told that the pad lanes are its own
        (mi_set_pad_lanes_scratch): a slice of somebody else's tensor with the same strides never qualifies."""
        if buf.shape[-1] == c:
            return buf
        key = buf.data_ptr()
        self._padded[key] = (buf.shape[-1], c)
        weakref.finalize(buf, self._padded.pop, key, None)
        return buf[..., :c]

    def _owns_pad_lanes(self, t):
        rec = self._padded.get(t.data_ptr())
        return rec is not None and rec == (t.stride(-2), t.shape[-1]) and t.storage_offset() == 0

    def empty_like_act(self, t):
        n, h, w, c = t.shape
        return self.empty_act(n, h, w, c)

    def empty_weight(self, cout, cin, k):
        ld = pad4(cin)
        buf = torch.zeros(cout, k, k, ld, device=self.device, dtype=torch.float32)
        return buf[..., :cin] if ld != cin else buf

    def set_workspace_slot(self, slot):
        """Split-K scratch is per execution lane: two tasks running on two streams must not share it."""
        self._ws_slot = slot

    def workspace(self, nbytes):
        if self._ws is None:
            self._ws = {}
        slot = getattr(self, "_ws_slot", 0)
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                # a captured CUDA graph may hold the old address: outgrown buffers are retired, never freed
                self._ws.setdefault("retired", []).append(ws)
            ws = torch.empty(int(nbytes), device=self.device, dtype=torch.uint8)
            if _POISON:
                ws.fill_(255)          # 0xFFFFFFFF is a NaN pattern
            self._ws[slot] = ws
        return ws

    def launch_count(self):
        """libmi_b200 kernels executed so far: eager launches + kernels inside replayed CUDA graphs."""
        return int(self.lib.mi_launch_count()) + self.replayed_launches

    # ------------------------------------------------------------------ per-launch profiling (bench roofline)
    PROF_TAGS = {"fprop_tc": 0, "wgrad_tc": 1, "fprop_simt": 2, "wgrad_simt": 3, "sepconv_fwd": 4, "sepconv_bwd": 5,
                 "wgrad_finish": 6, "fprop_tc_halo": 7, "fprop_tc_halo_stream": 8, "wgrad_tc_kx": 9, "fprop_tc_kxs": 10}

    def prof_enable(self, on):
        _lib.check(self.lib.mi_prof_enable(1 if on else 0), "mi_prof_enable")

    def prof_summary(self):
        """{tag: dict(launches, ms, flops, bytes)} since prof_enable(True) (synchronises the device)."""
        import ctypes
        out = {}
        buf = (ctypes.c_double * 4)()
        for name, tag in self.PROF_TAGS.items():
            _lib.check(self.lib.mi_prof_summary(tag, ctypes.cast(buf, ctypes.c_void_p)), "mi_prof_summary")
            out[name] = dict(launches=int(buf[0]), ms=buf[1], flops=buf[2], bytes=buf[3])
        return out

    # ------------------------------------------------------------------ convolution
    def conv_fprop(self, x, w, b, act=ACT_NONE, slope=0.0, out=None, engine=None):
        n, h, wd, cin = x.shape
        cout, k = w.shape[0], w.shape[1]
        assert w.shape[3] == cin
        y = out if out is not None else self.empty_act(n, h, wd, cout)
        scratch = (cout & 3) and self._owns_pad_lanes(y)
        if scratch:
            self.lib.mi_set_pad_lanes_scratch(1)
        try:
            rc = self.lib.mi_conv2d_fprop(x.data_ptr(), _ld(x), w.data_ptr(), _ldw(w), self._p(b), y.data_ptr(),
                                          _ld(y), n, h, wd, cin, cout, k, act, float(slope),
                                          self.engine if engine is None else engine, self._stream())
        finally:
            if scratch:
                self.lib.mi_set_pad_lanes_scratch(0)
        _lib.check(rc, "mi_conv2d_fprop")
        return y

    def set_sm_budget(self, ctas):
        """CTAs a persistent tensor-core launch may occupy from now on (0 = every SM); returns the value in force."""
        return int(self.lib.mi_set_sm_budget(int(ctas)))

    def weight_to_dgrad(self, w, out=None, rnd=None):
        """Rotated / transposed copy of a KRSC weight for dgrad; rounded to the TF32 grid under ``tf32_rn``."""
        cout, k, _, cin = w.shape
        wt = out if out is not None else self.empty_weight(cin, cout, k)
        rnd = self.tf32_rn if rnd is None else rnd
        _lib.check(self.lib.mi_weight_to_dgrad(w.data_ptr(), _ldw(w), wt.data_ptr(), _ldw(wt), cin, cout, k,
                                               1 if rnd else 0, self._stream()), "mi_weight_to_dgrad")
        return wt

    def round_tf32(self, x, out=None):
        """Round-to-nearest onto the TF32 grid: an NHWC activation / KRSC weight view (rows of the last dim, any row
        stride) or a flat contiguous buffer; in place unless ``out`` is given."""
        y = x if out is None else out
        if x.dim() == 1:
            c, rows, ldx, ldy = 4, x.numel() // 4, 4, 4
            assert x.numel() % 4 == 0 and x.is_contiguous() and y.is_contiguous()
        else:
            c = x.shape[-1]
            rows = x.numel() // c
            ldx, ldy = _ld(x), _ld(y)
        _lib.check(self.lib.mi_round_tf32(x.data_ptr(), ldx, y.data_ptr(), ldy, c, rows, self._stream()),
                   "mi_round_tf32")
        return y

    def conv_dgrad(self, dy, w, wt=None, mask_y=None, mask_act=ACT_NONE, mask_slope=0.0, out=None, accumulate=False,
                   engine=None):
        n, h, wd, cout = dy.shape
        k, cin = w.shape[1], w.shape[3]
        if wt is None:
            eng = self.engine if engine is None else engine
            wt = self.weight_to_dgrad(w, rnd=self.tf32_rn and eng != ENGINE_SIMT)
        dx = out if out is not None else self.empty_act(n, h, wd, cin)
        scratch = (cin & 3) and self._owns_pad_lanes(dx)
        if scratch:
            self.lib.mi_set_pad_lanes_scratch(1)
        try:
            rc = self.lib.mi_conv2d_dgrad(dy.data_ptr(), _ld(dy), wt.data_ptr(), _ldw(wt), dx.data_ptr(), _ld(dx),
                                          self._p(mask_y), 0 if mask_y is None else _ld(mask_y), mask_act,
                                          float(mask_slope), 1 if accumulate else 0, n, h, wd, cin, cout, k,
                                          self.engine if engine is None else engine, self._stream())
        finally:
            if scratch:
                self.lib.mi_set_pad_lanes_scratch(0)
        _lib.check(rc, "mi_conv2d_dgrad")
        return dx

    def begin_deferred_wgrad(self):
        """From here to ``flush_deferred_wgrad`` the finishing stage of every ``conv_wgrad`` (reduction of the split-K
        partials, bias, store / accumulate / fused update) is recorded and then issued as one launch per 20 layers."""
        _lib.check(self.lib.mi_wgrad_defer_begin(), "mi_wgrad_defer_begin")
        self._deferred_ws = []

    def flush_deferred_wgrad(self):
        held, self._deferred_ws = self._deferred_ws, None
        _lib.check(self.lib.mi_wgrad_defer_flush(self._stream()), "mi_wgrad_defer_flush")
        del held           # the partials of each layer had to stay alive (and distinct) until this point

    def conv_wgrad(self, x, dy, k, ldw, spec, engine=None):
        n, h, wd, cin = x.shape
        cout = dy.shape[3]
        eng = self.engine if engine is None else engine
        need = self.lib.mi_conv2d_wgrad_workspace(n, h, wd, cin, cout, k, eng)
        if getattr(self, "_deferred_ws", None) is not None:
            ws = torch.empty(int(need), device=self.device, dtype=torch.uint8)
            if _POISON:
                ws.fill_(255)
            self._deferred_ws.append(ws)
        else:
            ws = self.workspace(need)
        p = self._p
        _lib.check(self.lib.mi_conv2d_wgrad(x.data_ptr(), _ld(x), dy.data_ptr(), _ld(dy), n, h, wd, cin, cout, k, ldw,
                                            spec.mode, float(spec.scale), p(spec.grad_w), p(spec.grad_b), p(spec.w_in),
                                            p(spec.b_in), p(spec.w_out), p(spec.b_out), p(spec.lr_w), p(spec.lr_b),
                                            p(spec.gsum_w), p(spec.gsum_b), p(spec.wt_out),
                                            0 if spec.wt_out is None else _ldw(spec.wt_out), p(spec.wr_out),
                                            ws.data_ptr(),
                                            ws.numel(), eng, self._stream()), "mi_conv2d_wgrad")

    # ------------------------------------------------------------------ resampling / pointwise
    def avgpool_fwd(self, x, rnd=False):
        n, h, w, c = x.shape
        y = self.empty_act(n, h // 2, w // 2, c)
        _lib.check(self.lib.mi_avgpool2_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), n, h, w, c, int(rnd),
                                            self._stream()), "mi_avgpool2_fwd")
        return y

    def avgpool_bwd(self, dy, dx, accumulate):
        n, h, w, c = dx.shape
        _lib.check(self.lib.mi_avgpool2_bwd(dy.data_ptr(), _ld(dy), dx.data_ptr(), _ld(dx), int(accumulate), n, h, w,
                                            c, self._stream()), "mi_avgpool2_bwd")

    def maxpool_fwd(self, x):
        n, h, w, c = x.shape
        y = self.empty_act(n, h // 2, w // 2, c)
        _lib.check(self.lib.mi_maxpool2_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), n, h, w, c, self._stream()),
                   "mi_maxpool2_fwd")
        return y

    def maxpool_bwd(self, x, dy, dx, accumulate):
        n, h, w, c = x.shape
        _lib.check(self.lib.mi_maxpool2_bwd(x.data_ptr(), _ld(x), dy.data_ptr(), _ld(dy), dx.data_ptr(), _ld(dx),
                                            int(accumulate), n, h, w, c, self._stream()), "mi_maxpool2_bwd")

    def upsample_fwd(self, x, align_corners, out=None, rnd=False):
        n, h, w, c = x.shape
        y = out if out is not None else self.empty_act(n, 2 * h, 2 * w, c)
        _lib.check(self.lib.mi_upsample2_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), n, h, w, c,
                                             int(align_corners), int(rnd), self._stream()), "mi_upsample2_fwd")
        return y

    def upsample_bwd(self, dy, dx, align_corners, accumulate, mask_y=None, mask_act=ACT_NONE, mask_slope=0.0,
                     rnd=False):
        n, h, w, c = dx.shape
        if mask_y is not None:   # fused activation derivative: the plain upsample is the window that covers everything
            return self.upsample_window_bwd(dy, dx, align_corners, accumulate, (h, w), (0, 0), (0, 0), mask_y,
                                            mask_act, mask_slope, rnd=rnd)
        _lib.check(self.lib.mi_upsample2_bwd(dy.data_ptr(), _ld(dy), dx.data_ptr(), _ld(dx), int(accumulate), n, h, w,
                                             c, int(align_corners), int(rnd), self._stream()), "mi_upsample2_bwd")

    def upsample_window_fwd(self, x, align_corners, full_hw, lo_origin, hi_origin, hi_hw, rnd=False):
        """x = rows/cols [lo_origin, +x.shape) of a full_hw grid -> rows/cols [hi_origin, +hi_hw) of its x2 upsampling."""
        n, h, w, c = x.shape
        y = self.empty_act(n, hi_hw[0], hi_hw[1], c)
        _lib.check(self.lib.mi_upsample2_window_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), n, h, w, c,
                                                    int(align_corners), full_hw[0], full_hw[1], lo_origin[0],
                                                    lo_origin[1], hi_hw[0], hi_hw[1], hi_origin[0], hi_origin[1],
                                                    int(rnd), self._stream()), "mi_upsample2_window_fwd")
        return y

    def upsample_window_bwd(self, dy, dx, align_corners, accumulate, full_hw, lo_origin, hi_origin, mask_y=None,
                            mask_act=ACT_NONE, mask_slope=0.0, rnd=False):
        n, h, w, c = dx.shape
        _lib.check(self.lib.mi_upsample2_window_bwd(dy.data_ptr(), _ld(dy), dx.data_ptr(), _ld(dx), int(accumulate), n,
                                                    h, w, c, int(align_corners), full_hw[0], full_hw[1], lo_origin[0],
                                                    lo_origin[1], dy.shape[1], dy.shape[2], hi_origin[0], hi_origin[1],
                                                    self._p(mask_y), 0 if mask_y is None else _ld(mask_y), mask_act,
                                                    float(mask_slope), int(rnd), self._stream()),
                   "mi_upsample2_window_bwd")

    def window_copy(self, src, src_origin, dst, dst_origin, hw, accumulate=False):
        """dst[:, dy0:dy0+h, dx0:dx0+w] (+)= src[:, sy0:sy0+h, sx0:sx0+w] between two NHWC buffers."""
        n, sh, sw, c = src.shape
        _, dh, dw, _ = dst.shape
        _lib.check(self.lib.mi_window_copy(src.data_ptr(), _ld(src), sh, sw, src_origin[0], src_origin[1],
                                           dst.data_ptr(), _ld(dst), dh, dw, dst_origin[0], dst_origin[1], n, hw[0],
                                           hw[1], c, int(accumulate), self._stream()), "mi_window_copy")

    def add(self, a, b, out=None, rnd=False):
        n, h, w, c = a.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_add(a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), y.data_ptr(), _ld(y), n * h * w, c,
                                   int(rnd), self._stream()), "mi_add")
        return y

    def copy(self, src, dst, accumulate=False):
        n, h, w, c = src.shape
        _lib.check(self.lib.mi_copy(src.data_ptr(), _ld(src), dst.data_ptr(), _ld(dst), int(accumulate), n * h * w, c,
                                    self._stream()), "mi_copy")

    def act_bwd(self, dy, y, act, slope=0.0, rnd=False):
        n, h, w, c = y.shape
        _lib.check(self.lib.mi_act_bwd(dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), act, float(slope), n * h * w, c,
                                       1 if rnd else 0, self._stream()), "mi_act_bwd")

    def fill(self, t, value):
        assert t.is_contiguous()
        _lib.check(self.lib.mi_fill(t.data_ptr(), float(value), t.numel(), self._stream()), "mi_fill")

    def axpby(self, x, a, y, b):
        assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
        _lib.check(self.lib.mi_axpby(x.data_ptr(), float(a), y.data_ptr(), float(b), x.numel(), self._stream()),
                   "mi_axpby")

    # ------------------------------------------------------------------ glue ops of the flow-based backbones
    BIN_ADD, BIN_SUB, BIN_MUL, BIN_DIV = 0, 1, 2, 3

    def bn_eval_fwd(self, x, gamma, beta, mean, var, eps, act=ACT_NONE, slope=0.0, out=None):
        n, h, w, c = x.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_bn_eval_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), gamma.data_ptr(),
                                           beta.data_ptr(), mean.data_ptr(), var.data_ptr(), float(eps), act,
                                           float(slope), n * h * w, c, self._stream()), "mi_bn_eval_fwd")
        return y

    def bn_eval_bwd(self, dy, y, x, gamma, mean, var, eps, act, slope, dx, accumulate_dx, dgamma, dbeta, mode, scale):
        n, h, w, c = x.shape
        ws = self.workspace(self.lib.mi_bn_eval_bwd_workspace(n * h * w, c))
        _lib.check(self.lib.mi_bn_eval_bwd(dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), x.data_ptr(), _ld(x),
                                           self._p(dx), 0 if dx is None else _ld(dx), int(accumulate_dx),
                                           gamma.data_ptr(), mean.data_ptr(), var.data_ptr(), float(eps), act,
                                           float(slope), self._p(dgamma), self._p(dbeta), mode, float(scale),
                                           ws.data_ptr(), ws.numel(), n * h * w, c, self._stream()), "mi_bn_eval_bwd")

    def binary_fwd(self, op, a, b, out=None):
        n, h, w, c = a.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_binary_fwd(op, a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), b.shape[3], y.data_ptr(),
                                          _ld(y), n * h * w, c, self._stream()), "mi_binary_fwd")
        return y

    def binary_bwd(self, op, a, b, go, ga, acc_a, gb, acc_b):
        n, h, w, c = a.shape
        _lib.check(self.lib.mi_binary_bwd(op, a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), b.shape[3], go.data_ptr(),
                                          _ld(go), self._p(ga), 0 if ga is None else _ld(ga), int(acc_a), self._p(gb),
                                          0 if gb is None else _ld(gb), int(acc_b), n * h * w, c, self._stream()),
                   "mi_binary_bwd")

    def affine(self, x, alpha, beta, out=None, accumulate=False):
        n, h, w, c = x.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_affine(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), float(alpha), float(beta),
                                      int(accumulate), n * h * w, c, self._stream()), "mi_affine")
        return y

    # ------------------------------------------------------------------ heads of the flow / attention backbones
    BLEND_RATIO, BLEND_RATIO_COMPLEMENT, BLEND_LERP = 0, 1, 2
    RING_ZERO, RING_REFLECT = 0, 1

    def act_fwd(self, x, act, slope=0.0, out=None):
        n, h, w, c = x.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_act_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), act, float(slope), n * h * w, c,
                                       self._stream()), "mi_act_fwd")
        return y

    def clamp_fwd(self, x, lo, hi, out=None):
        n, h, w, c = x.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_clamp_fwd(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), float(lo), float(hi), n * h * w,
                                         c, self._stream()), "mi_clamp_fwd")
        return y

    def clamp_bwd(self, dy, x, dx, lo, hi, accumulate):
        n, h, w, c = x.shape
        _lib.check(self.lib.mi_clamp_bwd(dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), dx.data_ptr(), _ld(dx),
                                         int(accumulate), float(lo), float(hi), n * h * w, c, self._stream()),
                   "mi_clamp_bwd")

    def blend_fwd(self, a, b, m0, m1, w0, w1, eps, mode, out=None):
        n, h, w, c = a.shape
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_blend_fwd(a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), m0.data_ptr(), _ld(m0),
                                         self._p(m1), 0 if m1 is None else _ld(m1), y.data_ptr(), _ld(y), float(w0),
                                         float(w1), float(eps), mode, n * h * w, c, self._stream()), "mi_blend_fwd")
        return y

    def blend_bwd(self, a, b, m0, m1, go, ga, gb, gm0, gm1, accumulate, w0, w1, eps, mode):
        n, h, w, c = a.shape
        ld = lambda t: 0 if t is None else _ld(t)
        _lib.check(self.lib.mi_blend_bwd(a.data_ptr(), _ld(a), b.data_ptr(), _ld(b), m0.data_ptr(), _ld(m0),
                                         self._p(m1), ld(m1), go.data_ptr(), _ld(go), self._p(ga), ld(ga), self._p(gb),
                                         ld(gb), self._p(gm0), ld(gm0), self._p(gm1), ld(gm1), int(accumulate),
                                         float(w0), float(w1), float(eps), mode, n * h * w, c, self._stream()),
                   "mi_blend_bwd")

    def ring_fix(self, x, mode):
        n, h, w, c = x.shape
        _lib.check(self.lib.mi_ring_fix(x.data_ptr(), _ld(x), n, h, w, c, mode, self._stream()), "mi_ring_fix")

    def ring_fold(self, g, mode):
        n, h, w, c = g.shape
        _lib.check(self.lib.mi_ring_fold(g.data_ptr(), _ld(g), n, h, w, c, mode, self._stream()), "mi_ring_fold")

    def channel_mean_nchw(self, f):
        n, c, h, w = f.shape
        assert f.is_contiguous()
        out = torch.empty(n * c, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.mi_channel_mean_nchw(f.data_ptr(), out.data_ptr(), n * c, h * w, self._stream()),
                   "mi_channel_mean_nchw")
        return out

    def space_to_depth(self, f0, f1, mean0, mean1, pad_top, pad_left, oh, ow, r):
        """Two NCHW frames -> one ringed NHWC buffer [n, oh+2, ow+2, 6*r*r] (reflect pad, mean shift, ring 0)."""
        n, c, h, w = f0.shape
        assert c == 3 and f0.is_contiguous() and f1.is_contiguous()
        out = self.empty_act(n, oh + 2, ow + 2, 6 * r * r)
        _lib.check(self.lib.mi_space_to_depth(f0.data_ptr(), f1.data_ptr(), mean0.data_ptr(), mean1.data_ptr(),
                                              out.data_ptr(), _ld(out), n, h, w, pad_top, pad_left, oh, ow, r,
                                              self._stream()), "mi_space_to_depth")
        return out

    def depth_to_space(self, x, mean0, mean1, h, w, pad_top, pad_left, r):
        """Ringed NHWC [n, ih+2, iw+2, 3*r*r] -> cropped NCHW [n,3,h,w] + (mean0+mean1)/2."""
        n, hh, ww, c = x.shape
        assert c == 3 * r * r
        out = torch.empty(n, 3, h, w, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.mi_depth_to_space(x.data_ptr(), _ld(x), mean0.data_ptr(), mean1.data_ptr(), out.data_ptr(),
                                              n, h, w, pad_top, pad_left, hh - 2, ww - 2, r, self._stream()),
                   "mi_depth_to_space")
        return out

    def depth_to_space_bwd(self, gout, gin, pad_top, pad_left, r):
        n, _, h, w = gout.shape
        _, hh, ww, c = gin.shape
        assert gout.is_contiguous()
        _lib.check(self.lib.mi_depth_to_space_bwd(gout.data_ptr(), gin.data_ptr(), _ld(gin), n, h, w, pad_top,
                                                  pad_left, hh - 2, ww - 2, r, self._stream()),
                   "mi_depth_to_space_bwd")

    def interior_reduce(self, x, mul, ring, scale):
        """[n,h,w,c] -> [n,1,1,c]: scale * sum over the interior (ring excluded) of x (* mul)."""
        n, h, w, c = x.shape
        out = self.empty_act(n, 1, 1, c)
        assert _ld(out) == c
        _lib.check(self.lib.mi_interior_reduce(x.data_ptr(), _ld(x), self._p(mul), 0 if mul is None else _ld(mul),
                                               out.data_ptr(), n, h, w, c, ring, float(scale), self._stream()),
                   "mi_interior_reduce")
        return out

    def scale_add(self, o, s, res, out=None):
        n, h, w, c = o.shape
        assert s.shape == (n, 1, 1, c) and _ld(s) == c
        y = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_scale_add(o.data_ptr(), _ld(o), s.data_ptr(), self._p(res),
                                         0 if res is None else _ld(res), y.data_ptr(), _ld(y), n, h * w, c,
                                         self._stream()), "mi_scale_add")
        return y

    def scale_bwd(self, g, s, dx, accumulate):
        n, h, w, c = g.shape
        assert _ld(s) == c
        _lib.check(self.lib.mi_scale_bwd(g.data_ptr(), _ld(g), s.data_ptr(), dx.data_ptr(), _ld(dx), int(accumulate),
                                         n, h * w, c, self._stream()), "mi_scale_bwd")

    def interior_bcast_add(self, dy, dx, ring, scale):
        n, h, w, c = dx.shape
        assert _ld(dy) == c
        _lib.check(self.lib.mi_interior_bcast_add(dy.data_ptr(), dx.data_ptr(), _ld(dx), n, h, w, c, ring,
                                                  float(scale), self._stream()), "mi_interior_bcast_add")

    # ------------------------------------------------------------------ frames in / prediction out
    def frames_to_canvas(self, f0, f1, ch, cw, pad_top, pad_left, mode, rnd=False):
        """f0, f1: NCHW [n,3,h,w] contiguous -> NHWC canvas [n,ch,cw,6] (ld 8)."""
        n, c, h, w = f0.shape
        assert c == 3 and f0.is_contiguous() and f1.is_contiguous()
        y = self.empty_act(n, ch, cw, 6)
        _lib.check(self.lib.mi_frames_to_canvas(f0.data_ptr(), f1.data_ptr(), y.data_ptr(), _ld(y), n, h, w, ch, cw,
                                                pad_top, pad_left, mode, int(rnd), self._stream()),
                   "mi_frames_to_canvas")
        return y

    def nhwc_window_to_nchw(self, src, y0, x0, h, w):
        n, hs, ws, c = src.shape
        dst = torch.empty(n, c, h, w, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.mi_nhwc_window_to_nchw(src.data_ptr(), _ld(src), dst.data_ptr(), n, hs, ws, y0, x0, h, w,
                                                   c, self._stream()), "mi_nhwc_window_to_nchw")
        return dst

    def nchw_to_nhwc_window(self, src, dst, y0, x0):
        n, c, h, w = src.shape
        _, hs, ws, _ = dst.shape
        assert src.is_contiguous()
        _lib.check(self.lib.mi_nchw_to_nhwc_window(src.data_ptr(), dst.data_ptr(), _ld(dst), n, hs, ws, y0, x0, h, w,
                                                   c, self._stream()), "mi_nchw_to_nhwc_window")

    # ------------------------------------------------------------------ adaptive separable convolution
    def sepconv_planar(self, n, oh, ow, taps):
        """Workspace for the tap-planar copies of one sepconv call's two filter tensors (see include/mi_b200.h)."""
        nbytes = int(self.lib.mi_sepconv_planar_bytes(n, oh, ow, taps))
        return torch.empty(nbytes // 4, device=self.device, dtype=torch.float32)

    def sepconv_fwd(self, frame, vert, horiz, oh, ow, gy0, gx0, iy0, ix0, planar=None):
        """frame NCHW [n,c,fh,fw]; vert/horiz NHWC [n,gh,gw,F]; -> NCHW [n,c,oh,ow].  ``planar`` (from
        ``sepconv_planar``) selects the four-pixels-per-thread kernels and receives the planar filters."""
        n, c, fh, fw = frame.shape
        _, gh, gw, taps = vert.shape
        assert frame.is_contiguous() and _ld(vert) == _ld(horiz)
        out = torch.empty(n, c, oh, ow, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.mi_sepconv_fwd(frame.data_ptr(), vert.data_ptr(), horiz.data_ptr(), _ld(vert),
                                           out.data_ptr(), n, c, fh, fw, gh, gw, oh, ow, gy0, gx0, iy0, ix0, taps,
                                           self._p(planar), self._stream()), "mi_sepconv_fwd")
        return out

    def sepconv_bwd(self, frame, vert, horiz, grad_out, g_vert, g_horiz, gy0, gx0, iy0, ix0, rnd=False, planar=None,
                    planar_valid=False, planar_grad=None, zero_outside=False):
        """``zero_outside``: g_vert / g_horiz are fresh (uninitialised) buffers and the call zero-fills what lies outside
        the window itself (MI_SEPCONV_ZERO_OUTSIDE); otherwise the caller has zero-filled them."""
        n, c, fh, fw = frame.shape
        _, gh, gw, taps = vert.shape
        oh, ow = grad_out.shape[2], grad_out.shape[3]
        assert grad_out.is_contiguous() and _ld(g_vert) == _ld(g_horiz) and _ld(vert) == _ld(horiz)
        _lib.check(self.lib.mi_sepconv_bwd(frame.data_ptr(), vert.data_ptr(), horiz.data_ptr(), _ld(vert),
                                           grad_out.data_ptr(), g_vert.data_ptr(), g_horiz.data_ptr(), _ld(g_vert), n,
                                           c, fh, fw, gh, gw, oh, ow, gy0, gx0, iy0, ix0, taps,
                                           int(bool(rnd)) | (2 if zero_outside else 0), self._p(planar),
                                           int(planar_valid), self._p(planar_grad), self._stream()),
                   "mi_sepconv_bwd")

    # ------------------------------------------------------------------ warp
    def warp_fwd(self, img, flow, variant, sx=1.0, sy=1.0, out=None):
        n, h, w, c = img.shape
        out = out if out is not None else self.empty_act(n, h, w, c)
        _lib.check(self.lib.mi_warp_fwd(img.data_ptr(), _ld(img), flow.data_ptr(), _ld(flow), out.data_ptr(), _ld(out),
                                        n, h, w, c, variant, float(sx), float(sy), self._stream()), "mi_warp_fwd")
        return out

    def warp_bwd(self, img, flow, grad_out, grad_flow, variant, sx=1.0, sy=1.0, accumulate=False):
        n, h, w, c = img.shape
        _lib.check(self.lib.mi_warp_bwd(img.data_ptr(), _ld(img), flow.data_ptr(), _ld(flow), grad_out.data_ptr(),
                                        _ld(grad_out), grad_flow.data_ptr(), _ld(grad_flow), None, 0, int(accumulate),
                                        n, h, w, c, variant, float(sx), float(sy), self._stream()), "mi_warp_bwd")

    # ------------------------------------------------------------------ loss / metrics / optimizers
    def loss_fwd_bwd(self, pred, target, kind, weight, loss_out, grad=None):
        """loss_out[0] += weight*mean(f(pred-target)); grad (optional) = d/dpred."""
        assert pred.is_contiguous() and target.is_contiguous() and pred.numel() == target.numel()
        _lib.check(self.lib.mi_loss_fwd_bwd(pred.data_ptr(), target.data_ptr(), self._p(grad), loss_out.data_ptr(),
                                            pred.numel(), kind, float(weight), self._stream()), "mi_loss_fwd_bwd")

    def psnr_accumulate(self, pred, target, sq_out):
        assert pred.is_contiguous() and target.is_contiguous() and sq_out.dtype == torch.float64
        _lib.check(self.lib.mi_psnr_accumulate(pred.data_ptr(), target.data_ptr(), sq_out.data_ptr(), pred.numel(),
                                               self._stream()), "mi_psnr_accumulate")

    def ssim_accumulate(self, pred, target, window, sum_out, val_range=255.0):
        """sum_out[0] (double) += sum of the SSIM map of two [c,h,w] images in [0,1] (8-bit quantised in the kernel);
        ``window`` = normalised 1-D Gaussian (CPU float32 tensor, <= 11 taps)."""
        c, h, w = pred.shape
        assert pred.is_contiguous() and target.is_contiguous() and sum_out.dtype == torch.float64
        assert window.device.type == "cpu" and window.dtype == torch.float32 and window.is_contiguous()
        _lib.check(self.lib.mi_ssim_accumulate(pred.data_ptr(), target.data_ptr(), sum_out.data_ptr(), c, h, w,
                                               window.data_ptr(), window.numel(), float(val_range), self._stream()),
                   "mi_ssim_accumulate")

    def septuplet_prepare(self, src, y0, x0, reversed_, h, w, bgr=True, div255=True, mean=None, std=None):
        """Decoded uint8 frames [tasks,frames,H,W,3] -> float [frames,tasks,3,h,w]: crop, temporal flip, channel order,
        /255 and normalisation of data/vimeo_septuplet.py:50-78 in one launch."""
        import ctypes as C
        assert src.dtype == torch.uint8 and src.is_cuda and src.is_contiguous() and src.dim() == 5 and src.shape[4] == 3
        tasks, frames, sh, sw, _ = src.shape
        assert y0.dtype == torch.int32 and x0.dtype == torch.int32 and reversed_.dtype == torch.uint8
        assert y0.numel() == tasks and x0.numel() == tasks and reversed_.numel() == tasks
        assert (mean is None) == (std is None)
        out = torch.empty(frames, tasks, 3, h, w, device=src.device, dtype=torch.float32)
        m = (C.c_float * 3)(*[float(v) for v in mean]) if mean is not None else None
        s = (C.c_float * 3)(*[float(v) for v in std]) if std is not None else None
        _lib.check(self.lib.mi_septuplet_prepare(src.data_ptr(), out.data_ptr(), y0.data_ptr(), x0.data_ptr(),
                                                 reversed_.data_ptr(), tasks, frames, sh, sw, int(h), int(w), int(bgr),
                                                 int(div255), C.cast(m, C.c_void_p) if m is not None else None,
                                                 C.cast(s, C.c_void_p) if s is not None else None, self._stream()),
                   "mi_septuplet_prepare")
        return out

    def inner_update(self, w_in, g, w_out, exp_avg, exp_avg_sq, lr, lr_per_element, lr_stride, num_step, seg, skip,
                     rule, step_count):
        _lib.check(self.lib.mi_inner_update(w_in.data_ptr(), g.data_ptr(), w_out.data_ptr(), self._p(exp_avg),
                                            self._p(exp_avg_sq), lr.data_ptr(), int(lr_per_element), int(lr_stride),
                                            int(num_step), seg.data_ptr(), self._p(skip), w_in.numel(), rule,
                                            int(step_count), self._stream()), "mi_inner_update")

    def outer_step(self, p, g, m, v, kind, lr, beta1, beta2, eps, weight_decay, step):
        _lib.check(self.lib.mi_outer_step(p.data_ptr(), g.data_ptr(), self._p(m), self._p(v), p.numel(), kind,
                                          float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                          int(step), self._stream()), "mi_outer_step")

    def addcmul(self, y, a, x1, x2):
        assert y.is_contiguous() and x1.is_contiguous() and x2.is_contiguous()
        _lib.check(self.lib.mi_addcmul(y.data_ptr(), float(a), x1.data_ptr(), x2.data_ptr(), y.numel(),
                                       self._stream()), "mi_addcmul")

    def segment_dot(self, a, b, seg, out):
        """out[t] += <a_t, b_t> per arena tensor; b=None sums a alone."""
        _lib.check(self.lib.mi_segment_dot(a.data_ptr(), self._p(b), seg.data_ptr(), out.data_ptr(), a.numel(),
                                           self._stream()), "mi_segment_dot")

    def segment_scale(self, x, scale, seg, mask, y, alpha=1.0, accumulate=False):
        """y (+)= alpha * s(t) * x with s(t) = scale[t] on the tensors selected by mask (None: all), 1 elsewhere."""
        assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
        _lib.check(self.lib.mi_segment_scale(x.data_ptr(), scale.data_ptr(), seg.data_ptr(), self._p(mask),
                                             y.data_ptr(), float(alpha), int(accumulate), x.numel(), self._stream()),
                   "mi_segment_scale")
