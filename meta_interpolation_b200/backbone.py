"""Common machinery of the model plugins (the reference's ``net`` objects).

A plugin keeps the reference's plugin API:
``net.forward(frame0, frame1, params=dict_or_None, **kwargs) -> Tensor[1,3,H,W]``,
``net.zero_grad(params=None)``, ``net.restore_backup_stats()``, and
``named_parameters()`` with the reference's names and OIHW shapes -- but its
parameters are views into one flat arena and its compute is a tape of sm_100a
kernels.  ``forward`` is a single ``torch.autograd.Function`` so the reference's
own calling pattern (``torch.autograd.grad(loss, fast_weights.values(),
allow_unused=True)``, meta_learning_system.py:291-292) keeps working, including
``None`` gradients for tensors the reference never routes.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops as ops_mod
from .arena import Arena, Layout, pad4
from .ops import WG_STORE, WgradSpec
from .tape import ConvParam, Tape, Var

_default_ops = None


def default_ops():
    """The process-wide CUDA operator table (raises when the library or the GPU is missing)."""
    global _default_ops
    if _default_ops is None:
        _default_ops = ops_mod.CudaOps()
    return _default_ops


def set_default_ops(ops):
    """Test hook: inject another operator table (the CPU suite injects the oracle's)."""
    global _default_ops
    _default_ops = ops


class _BnSlot(nn.Module):
    """Reference-visible parameters / buffers of one frozen BatchNorm2d (voxel_flow.py:241-263)."""

    def __init__(self, weight, bias, channels, device, eps=1e-5):
        super().__init__()
        self.weight = weight
        self.bias = bias
        self.eps = eps
        self.register_buffer("running_mean", torch.zeros(channels, device=device))
        self.register_buffer("running_var", torch.ones(channels, device=device))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long, device=device))


class _ConvSlot(nn.Module):
    """Holds the reference-visible ``weight`` / ``bias`` Parameters of one conv (views into the arena)."""

    def __init__(self, weight, bias):
        super().__init__()
        self.weight = weight
        if bias is not None:
            self.bias = bias


def kernel_weight_view(ops, t):
    """OIHW tensor -> KRSC kernel view; zero-copy when ``t`` already aliases KRSC storage."""
    kv = t.permute(0, 2, 3, 1)
    co, k, _, ci = kv.shape
    ld = kv.stride(2)
    ok = (kv.stride(3) == 1 and ld >= ci and ld % 4 == 0 and kv.stride(1) == k * ld and kv.stride(0) == k * k * ld
          and kv.data_ptr() % 16 == 0)
    if ok:
        return kv
    w = ops.empty_weight(co, ci, k)
    w.copy_(kv)   # layout conversion at the API boundary only (foreign OIHW tensors)
    return w


class StoreSink:
    """Weight-gradient sink of the compat path: plain gradients, only where autograd asks for them."""

    def __init__(self, ops, wanted, vec_like=None):
        self.ops = ops
        self.wanted = wanted          # set of parameter names whose gradient is needed
        self.grads = {}
        self.vec_like = vec_like      # name -> tensor shaped like a batch-norm scale

    def bn_targets(self, name):
        wn, bn = name + ".weight", name + ".bias"
        if wn not in self.wanted and bn not in self.wanted:
            return None
        ref = self.vec_like(name)
        gw, gb = torch.zeros_like(ref), torch.zeros_like(ref)
        self.grads[wn], self.grads[bn] = gw, gb
        return gw, gb, WG_STORE, 1.0

    def weight_grad(self, p, x, dy, k):
        wn, bn = p.name + ".weight", p.name + ".bias"
        if wn not in self.wanted and bn not in self.wanted:
            return
        cout, _, _, cin = p.w.shape
        gw = self.ops.empty_weight(cout, cin, k)
        gb = torch.zeros(cout, device=gw.device, dtype=gw.dtype) if p.b is not None else None
        self.ops.conv_wgrad(x, dy, k, gw.stride(2), WgradSpec(WG_STORE, grad_w=gw, grad_b=gb))
        self.grads[wn] = gw
        if gb is not None:
            self.grads[bn] = gb


def build_tape(net, frame0, frame1, names, tensors):
    """Tape forward of ``net`` with the given (possibly substituted) parameter tensors; returns (tape, output Var)."""
    ops = net.ops
    table = dict(zip(names, tensors))
    cache = {}

    def provider(name):
        p = cache.get(name)
        if p is None:
            w = kernel_weight_view(ops, table[name + ".weight"].detach())
            b = table.get(name + ".bias")
            p = ConvParam(name, w, None if b is None else b.detach().contiguous())
            cache[name] = p
        return p

    def vectors(name):
        slot = net.bn_slot(name)
        return (table[name + ".weight"].detach(), table[name + ".bias"].detach(), slot.running_mean,
                slot.running_var, slot.eps)

    tape = Tape(ops, provider, sink=None, vectors=vectors)
    out = net.build_graph(tape, frame0.detach().contiguous(), frame1.detach().contiguous())
    return tape, out


def collect_grads(net, tape, names, needs):
    """Tape backward with a StoreSink; gradients in the reference's OIHW form for the tensors autograd asks for."""
    wanted = {n for n, need in zip(names, needs) if need}
    sink = StoreSink(net.ops, wanted, vec_like=lambda name: net.bn_slot(name).weight.detach())
    tape.sink = sink
    tape.backward()
    grads = []
    for n, need in zip(names, needs):
        g = sink.grads.get(n) if need else None
        if g is not None and g.dim() == 4:
            g = g.permute(0, 3, 1, 2)
        grads.append(g)
    return grads


class _BackboneFunction(torch.autograd.Function):
    """forward: tape forward; backward: tape backward with a StoreSink."""

    @staticmethod
    def forward(ctx, net, frame0, frame1, names, *tensors):
        tape, out = build_tape(net, frame0, frame1, names, tensors)
        ctx.tape, ctx.out_var, ctx.names, ctx.net = tape, out, names, net
        ctx.shapes = [tuple(t.shape) for t in tensors]
        return out.data.clone() if net.clone_output else out.data

    @staticmethod
    def backward(ctx, grad_out):
        ctx.out_var.grad = grad_out.contiguous()
        grads = collect_grads(ctx.net, ctx.tape, ctx.names, ctx.needs_input_grad[4:])
        ctx.tape = ctx.out_var = None
        return (None, None, None, None) + tuple(grads)


class MetaBackbone(nn.Module):
    """Base of the five plugins.  Subclasses define ``conv_specs``, ``is_routed`` and ``build_graph``."""

    clone_output = False

    def __init__(self, ops=None):
        super().__init__()
        self.ops = ops if ops is not None else default_ops()

    # -- subclasses -----------------------------------------------------------------------------
    def conv_specs(self):
        """Ordered [(name, cin, cout, k, has_bias)] in the reference's registration order."""
        raise NotImplementedError

    def is_routed(self, param_name):
        """Does the reference feed this tensor from ``params``?  """
        return True

    def build_graph(self, tape, frame0, frame1):
        raise NotImplementedError

    # -- construction ---------------------------------------------------------------------------
    def param_entries(self):
        """Ordered entries in the reference's registration order: ("conv", name, cin, cout, k, has_bias) or
        ("bn", name, channels).  Default: the convs of ``conv_specs``."""
        return [("conv",) + tuple(spec) for spec in self.conv_specs()]

    def _build_parameters(self, init_fn):
        """Create the arena and the reference-named Parameters; ``init_fn(name, shape) -> CPU tensor`` is
        called in registration order so the CPU RNG stream matches the reference's constructors."""
        entries = self.param_entries()
        named_shapes = []
        for e in entries:
            if e[0] == "conv":
                _, name, cin, cout, k, has_bias = e
                named_shapes.append((name + ".weight", (cout, cin, k, k)))
                if has_bias:
                    named_shapes.append((name + ".bias", (cout,)))
            else:
                _, name, c = e
                named_shapes.append((name + ".weight", (c,)))
                named_shapes.append((name + ".bias", (c,)))
        self.layout = Layout(named_shapes)
        self.arena = Arena(self.layout, self.ops.device)
        self.param_names = [n for n, _ in named_shapes]
        self.conv_names = [e[1] for e in entries if e[0] == "conv"]
        self.bn_names = [e[1] for e in entries if e[0] == "bn"]
        self._spec = {e[1]: e[1:] for e in entries if e[0] == "conv"}
        self._bn_slots = {}
        for e in entries:
            if e[0] == "conv":
                _, name, cin, cout, k, has_bias = e
                wv = self.arena.reference_view(name + ".weight")
                wv.copy_(init_fn(name + ".weight", (cout, cin, k, k)))
                weight = nn.Parameter(wv)
                bias = None
                if has_bias:
                    bv = self.arena.reference_view(name + ".bias")
                    bv.copy_(init_fn(name + ".bias", (cout,)))
                    bias = nn.Parameter(bv)
                self._register(name, _ConvSlot(weight, bias))
            else:
                _, name, c = e
                wv = self.arena.reference_view(name + ".weight")
                wv.copy_(init_fn(name + ".weight", (c,)))
                bv = self.arena.reference_view(name + ".bias")
                bv.copy_(init_fn(name + ".bias", (c,)))
                slot = _BnSlot(nn.Parameter(wv), nn.Parameter(bv), c, self.ops.device)
                self._bn_slots[name] = slot
                self._register(name, slot)

    def bn_slot(self, name):
        return self._bn_slots[name]

    def meta_bn(self, name):
        slot = self._bn_slots[name]
        return (self.arena.kernel_view(name + ".weight"), self.arena.kernel_view(name + ".bias"), slot.running_mean,
                slot.running_var, slot.eps)

    def _register(self, dotted, slot):
        parts = dotted.split(".")
        mod = self
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        mod.add_module(parts[-1], slot)

    # -- parameter providers ----------------------------------------------------------------------
    def meta_param(self, name):
        w = self.arena.kernel_view(name + ".weight")
        b = self.arena.kernel_view(name + ".bias") if self._spec[name][4] else None
        return ConvParam(name, w, b)

    # -- reference plugin API ---------------------------------------------------------------------
    def resolve_params(self, params=None):
        """(names, tensors) the forward pass reads: ``params`` entries where the reference routes them, else own."""
        own = dict(self.named_parameters())
        names, tensors = [], []
        for n in self.param_names:
            t = own[n]
            if params is not None and n in params and self.is_routed(n):
                t = params[n]
            names.append(n)
            tensors.append(t)
        return tuple(names), tensors

    def forward(self, frame0, frame1, params=None, **kwargs):
        names, tensors = self.resolve_params(params)
        return _BackboneFunction.apply(self, frame0, frame1, names, *tensors)

    def zero_grad(self, params=None, set_to_none=True):
        # reference sepconv/model.py:352-367 (host-syncing prints dropped; semantics: clear .grad)
        if params is None:
            for p in self.parameters():
                p.grad = None
        else:
            for p in params.values():
                if p is not None and p.requires_grad and p.is_leaf:
                    p.grad = None

    def restore_backup_stats(self):
        pass  # no batch statistics on any of the five backbones
