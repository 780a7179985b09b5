"""Outer (meta) optimizer: one fused multi-tensor step per flat buffer.

Stands where the reference builds ``optim.Adam(betas=(0.9,0.99))`` / ``optim.Adamax``
/ ``optim.SGD`` over ``trainable_parameters()`` (meta_learning_system.py:132-143).
It IS a ``torch.optim.Optimizer`` (``ReduceLROnPlateau``, ``param_groups[...]['lr']``
and ``zero_grad`` keep working) but every parameter is a view into a flat buffer
with a matching flat gradient buffer, so ``step()`` is one ``mi_outer_step`` launch
per buffer and the NCCL all-reduce of the meta-gradient is one call per buffer.
"""
import torch

KIND = {"SGD": 0, "Adam": 1, "Adamax": 2}


class FlatGroup:
    """A flat parameter buffer, its flat gradient buffer and the Parameters viewing it."""

    def __init__(self, name, flat, members):
        """members: list of (Parameter, grad_view) where grad_view is the view of ``self.grad`` shaped like
        the Parameter (OIHW permuted view for conv weights)."""
        self.name = name
        self.flat = flat
        self.grad = torch.zeros_like(flat)
        self.members = members
        self.m = None
        self.v = None
        self.dirty = False     # fast path wrote gradients straight into ``self.grad``

    @staticmethod
    def pack(name, params, device):
        """Re-home arbitrary contiguous Parameters into one flat buffer (values preserved)."""
        total = sum(p.numel() for p in params)
        flat = torch.zeros(max(total, 1), device=device, dtype=torch.float32)
        group = FlatGroup(name, flat, [])
        off = 0
        for p in params:
            n = p.numel()
            view = flat[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            group.members.append((p, group.grad[off:off + n].view(p.shape)))
            off += n
        return group


class FusedOuterOptimizer(torch.optim.Optimizer):
    def __init__(self, groups, ops, kind, lr, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0):
        self.flat_groups = groups
        self.ops = ops
        self.kind = KIND[kind]
        params = [p for g in groups for p, _ in g.members]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0

    # ------------------------------------------------------------------ checkpoint interchange
    # The reference saves ``optimizer.state_dict()`` of a stock ``torch.optim.Adam / Adamax / SGD`` built over
    # ``trainable_parameters()`` (utils.py:34-75, experiment_builder.py save/load).  The moments here live in flat
    # buffers, so ``state_dict`` / ``load_state_dict`` translate to and from the stock per-parameter layout
    # (``state[i] = {step, exp_avg, exp_avg_sq | exp_inf}`` in the order of ``ref_order``).
    def set_reference_order(self, params, policies=None):
        """Parameter order of the reference's optimizer (``nn.Module.parameters()`` order of its system).
        ``policies``: the reference's list of param-group dicts when it builds its optimizer from several groups
        (voxelflow's ``get_optim_policies``); their extra keys and sizes shape ``state_dict()['param_groups']``."""
        own = {id(p) for g in self.flat_groups for p, _ in g.members}
        assert {id(p) for p in params} == own, "reference order must list exactly the optimised parameters"
        self._ref_order = list(params)
        self._ref_policies = None if policies is None else [
            ({k: v for k, v in g.items() if k != 'params'}, len(g['params'])) for g in policies]

    def _moment_views(self):
        """{id(param): (exp_avg view, second-moment view)} shaped like the parameter (views of the flat buffers)."""
        out = {}
        for g in self.flat_groups:
            for p, gv in g.members:
                off = gv.storage_offset() - g.grad.storage_offset()
                out[id(p)] = tuple(None if b is None else torch.as_strided(b, gv.size(), gv.stride(), off)
                                   for b in (g.m, g.v))
        return out

    def state_dict(self):
        order = getattr(self, "_ref_order", None) or [p for g in self.flat_groups for p, _ in g.members]
        views = self._moment_views()
        second = "exp_avg_sq" if self.kind == 1 else "exp_inf"
        state = {}
        if self.kind != 0 and self._step > 0:
            for i, p in enumerate(order):
                m, v = views[id(p)]
                state[i] = {"step": torch.tensor(float(self._step)), "exp_avg": m.detach().clone().contiguous(),
                            second: v.detach().clone().contiguous()}
        hp = dict(self.param_groups[0])
        policies = getattr(self, "_ref_policies", None)
        if policies is None:
            hp["params"] = list(range(len(order)))
            groups = [hp]
        else:
            groups, first = [], 0
            for extra, count in policies:
                g = dict(hp)
                g.update(extra)
                g["params"] = list(range(first, first + count))
                groups.append(g)
                first += count
        return {"state": state, "param_groups": groups, "fused_step": self._step}

    def load_state_dict(self, state_dict):
        order = getattr(self, "_ref_order", None) or [p for g in self.flat_groups for p, _ in g.members]
        hp_in = state_dict["param_groups"][0]
        for key in ("lr", "betas", "eps", "weight_decay"):
            if key in hp_in:
                self.param_groups[0][key] = hp_in[key]
        st = state_dict.get("state", {})
        if self.kind == 0 or not st:
            self._step = int(state_dict.get("fused_step", 0))
            return
        if len(st) != len(order):
            raise ValueError("optimizer state has %d entries for %d parameters" % (len(st), len(order)))
        for g in self.flat_groups:
            if g.m is None:
                g.m = torch.zeros_like(g.flat)
                g.v = torch.zeros_like(g.flat)
        views = self._moment_views()
        second = "exp_avg_sq" if self.kind == 1 else "exp_inf"
        steps = set()
        with torch.no_grad():
            for i, p in enumerate(order):
                e = st[i] if i in st else st[str(i)]
                m, v = views[id(p)]
                if tuple(e["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state %d has shape %s, parameter has %s"
                                     % (i, tuple(e["exp_avg"].shape), tuple(p.shape)))
                m.copy_(e["exp_avg"])
                v.copy_(e[second])
                steps.add(int(float(e["step"])))
        if len(steps) != 1:
            raise ValueError("per-parameter step counts differ: %s" % sorted(steps))
        self._step = steps.pop()

    def gather_grads(self):
        """Move autograd ``.grad`` tensors (compat path) into the flat gradient buffers."""
        for g in self.flat_groups:
            for p, gv in g.members:
                if p.grad is not None:
                    gv.add_(p.grad) if g.dirty else gv.copy_(p.grad)
                    p.grad = None

    def zero_grad(self, set_to_none=True):
        for g in self.flat_groups:
            self.ops.fill(g.grad, 0.0)
            g.dirty = False
            for p, _ in g.members:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        self.gather_grads()
        self._step += 1
        hp = self.param_groups[0]
        b1, b2 = hp["betas"]
        for g in self.flat_groups:
            if self.kind != 0 and g.m is None:
                g.m = torch.zeros_like(g.flat)
                g.v = torch.zeros_like(g.flat)
            self.ops.outer_step(g.flat, g.grad, g.m, g.v, self.kind, hp["lr"], b1, b2, hp["eps"],
                                hp["weight_decay"], self._step)
