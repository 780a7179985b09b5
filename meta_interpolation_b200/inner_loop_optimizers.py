"""Inner-loop learning rules (drop-in for the reference's ``inner_loop_optimizers.py``).

Same classes, constructor arguments, ``initialize`` / ``initialize_state`` /
``update_params(names_weights_dict, names_grads_wrt_params_dict, num_step, tau)``
contract and ``names_learning_rates_dict`` ParameterDict keys (``'.' -> '-'``) as
reference inner_loop_optimizers.py:57-244 (LSLR) and :248-426 (Meta-SGD).  The
update of each tensor is one fused kernel (``mi_inner_update``) instead of the
reference's chain of elementwise launches, and every reference quirk is kept
(SURVEY.md 3.4 / Appendix A):

* LSLR-SGD drops keys whose gradient is ``None`` (:141-142); Meta-SGD-SGD does not
  check and fails with the reference's TypeError (:329-330);
* LSLR-Adamax keeps ``exp_avg`` but loses ``exp_inf`` (:236); Meta-SGD-Adamax is
  fully stateless (:409, :418);
* moments live for the K steps of one task (``initialize_state`` per task, Q5).

``update_params`` is pure-functional like the reference: inputs are never
mutated and new tensors are returned; the returned tensors carry first-order
autograd history (identity w.r.t. the incoming weight, and the learning-rate
gradient of SURVEY Appendix E4), so the reference's ``loss.backward()`` over the
query pass yields the same meta-gradients.
"""
from collections import defaultdict

import torch
import torch.nn as nn

from .backbone import default_ops, kernel_weight_view

RULE_SGD, RULE_ADAM, RULE_ADAMAX_LSLR, RULE_ADAMAX_METASGD = 0, 1, 2, 3


def _storage_1d(t):
    """Flat view of the padded storage behind a kernel view (weights) or of a plain tensor."""
    if t.dim() == 4:
        co, k, _, ci = t.shape
        return t.as_strided((co * k * k * t.stride(2),), (1,), t.storage_offset())
    return t.reshape(-1) if t.is_contiguous() else None


def _kernel_form(ops, t):
    """Tensor in the layout the kernels use: KRSC-backed for 4-D weights, contiguous otherwise."""
    t = t.detach()
    if t.dim() == 4:
        return kernel_weight_view(ops, t)
    return t.contiguous()


def _alloc_like(ops, kv):
    if kv.dim() == 4:
        co, k, _, ci = kv.shape
        return ops.empty_weight(co, ci, k)
    return torch.zeros_like(kv)


def _reference_form(kv):
    return kv.permute(0, 3, 1, 2) if kv.dim() == 4 else kv


class _InnerUpdate(torch.autograd.Function):
    """w' = w - lr * direction(g, state): one fused kernel; first-order backward."""

    @staticmethod
    def forward(ctx, weight, lr, grad, rule, num_step, per_element, state, ops):
        kw = _kernel_form(ops, weight)
        kg = _kernel_form(ops, grad)
        out = _alloc_like(ops, kw)
        if rule in (RULE_ADAM, RULE_ADAMAX_LSLR, RULE_ADAMAX_METASGD):
            if len(state) == 0:
                state['step'] = 0
                state['exp_avg'] = _alloc_like(ops, kw)
                state['exp_avg_sq'] = _alloc_like(ops, kw)
            state['step'] += 1
        fw, fg, fo = _storage_1d(kw), _storage_1d(kg), _storage_1d(out)
        klr = _kernel_form(ops, lr) if per_element else lr.detach().contiguous()
        seg = torch.zeros((fw.numel() + 1023) // 1024, dtype=torch.int32, device=fw.device)
        ops.inner_update(fw, fg, fo,
                         _storage_1d(state['exp_avg']) if 'exp_avg' in state else None,
                         _storage_1d(state['exp_avg_sq']) if 'exp_avg_sq' in state else None,
                         _storage_1d(klr) if per_element else klr, per_element, 0, num_step, seg, None, rule,
                         state.get('step', 1))
        # SGD: the update direction IS the gradient; keep it instead of recovering it from w - w', which loses
        # most of its bits once lr * g drops towards the spacing of the weights (inner_lr defaults to 1e-5)
        ctx.save_for_backward(kw, kg if rule == RULE_SGD else out, klr)
        ctx.per_element, ctx.num_step, ctx.lr_shape, ctx.rule = per_element, num_step, tuple(lr.shape), rule
        return _reference_form(out)

    @staticmethod
    def backward(ctx, g_out):
        kw, second, klr = ctx.saved_tensors
        g_w = g_out if ctx.needs_input_grad[0] else None
        g_lr = None
        if ctx.needs_input_grad[1]:
            # update = w - lr*dir  =>  d/dlr = -<dir, G>  (SURVEY Appx E4)
            if ctx.rule == RULE_SGD:                     # dir = g (saved)
                direction = _reference_form(second)
                if ctx.per_element:
                    g_lr = -direction * g_out
                else:
                    g_lr = torch.zeros(ctx.lr_shape, device=g_out.device, dtype=g_out.dtype)
                    g_lr[ctx.num_step] = -(direction * g_out).sum()
            else:                                        # moment rules: dir recovered as (w - w')/lr
                rg = _reference_form(kw) - _reference_form(second)
                if ctx.per_element:
                    g_lr = -(rg / _reference_form(klr)) * g_out
                else:
                    g_lr = torch.zeros(ctx.lr_shape, device=g_out.device, dtype=g_out.dtype)
                    g_lr[ctx.num_step] = -(rg * g_out).sum() / klr[ctx.num_step]
        return g_w, g_lr, None, None, None, None, None, None


class _RuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.state = defaultdict(dict)
        self.beta1 = 0.9
        self.beta2 = 0.99
        self.weight_decay = 0
        self.eps = 1e-8
        self._ops = None

    @property
    def ops(self):
        if self._ops is None:
            self._ops = default_ops()
        return self._ops

    def initialize_state(self):
        # reference :104-105 / :293-294 -- called once per task, so moments span that task's K steps
        self.state = defaultdict(dict)


class LSLRGradientDescentLearningRule(_RuleBase):
    """Per-layer per-step learning rates (reference inner_loop_optimizers.py:57-244)."""

    def __init__(self, device, optimizer, total_num_inner_loop_steps, use_learnable_learning_rates,
                 init_learning_rate=1e-3):
        super().__init__()
        self.device = device
        self.init_learning_rate = torch.ones(1) * init_learning_rate
        self.total_num_inner_loop_steps = total_num_inner_loop_steps
        self.use_learnable_learning_rates = use_learnable_learning_rates
        self.optimizer = optimizer

    def initialize(self, names_weights_dict):
        self.names_learning_rates_dict = nn.ParameterDict()
        for idx, (key, param) in enumerate(names_weights_dict.items()):
            self.names_learning_rates_dict[key.replace(".", "-")] = nn.Parameter(
                data=(torch.ones(self.total_num_inner_loop_steps + 1) * self.init_learning_rate).to(param.device),
                requires_grad=self.use_learnable_learning_rates)

    def reset(self):
        pass

    def update_params(self, names_weights_dict, names_grads_wrt_params_dict, num_step, tau=0.1):
        if self.optimizer == 'SGD':
            rule = RULE_SGD
        elif self.optimizer == 'Adam':
            rule = RULE_ADAM
        elif self.optimizer == 'Adamax':
            rule = RULE_ADAMAX_LSLR
        else:
            raise NotImplementedError('This type of optimizer update operation is not yet implemented')
        updated = dict()
        for key in names_grads_wrt_params_dict.keys():
            g = names_grads_wrt_params_dict[key]
            if g is None:
                continue
            lr = self.names_learning_rates_dict[key.replace(".", "-")]
            updated[key] = _InnerUpdate.apply(names_weights_dict[key], lr, g, rule, num_step, False,
                                              self.state[key], self.ops)
        return updated

    # the reference exposes the three rules as methods too
    def update_sgd(self, names_weights_dict, names_grads_wrt_params_dict, num_step, tau=0.1):
        saved, self.optimizer = self.optimizer, 'SGD'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict, num_step, tau)
        finally:
            self.optimizer = saved

    def update_adam(self, names_weights_dict, names_grads_wrt_params_dict, num_step, tau=0.1, amsgrad=False):
        saved, self.optimizer = self.optimizer, 'Adam'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict, num_step, tau)
        finally:
            self.optimizer = saved

    def update_adamax(self, names_weights_dict, names_grads_wrt_params_dict, num_step, tau=0.1):
        saved, self.optimizer = self.optimizer, 'Adamax'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict, num_step, tau)
        finally:
            self.optimizer = saved


class MetaSGDLearningRule(_RuleBase):
    """Per-parameter learnable learning rates (reference inner_loop_optimizers.py:248-426)."""

    def __init__(self, device, optimizer, init_learning_rate=1e-3):
        super().__init__()
        assert init_learning_rate > 0., 'learning_rate should be positive.'
        self.init_learning_rate = init_learning_rate * torch.ones(1).to(device)
        self.device = device
        self.optimizer = optimizer

    def initialize(self, names_weights_dict):
        self.names_learning_rates_dict = nn.ParameterDict()
        for idx, (key, param) in enumerate(names_weights_dict.items()):
            if param.dim() == 4:   # alpha shares the KRSC-backed layout of its weight
                co, ci, k, _ = param.shape
                kv = self.ops.empty_weight(co, ci, k)
                kv.fill_(float(self.init_learning_rate))
                alpha = kv.permute(0, 3, 1, 2)
            else:
                alpha = torch.ones_like(param.detach()) * self.init_learning_rate.to(param.device)
            self.names_learning_rates_dict[key.replace(".", "-")] = nn.Parameter(alpha, requires_grad=True)

    def reset(self):
        for key, param in self.names_learning_rates_dict.items():
            param.data.fill_(float(self.init_learning_rate))

    def update_params(self, names_weights_dict, names_grads_wrt_params_dict, num_step=0, tau=0.1):
        if self.optimizer == 'SGD':
            rule = RULE_SGD
        elif self.optimizer == 'Adam':
            rule = RULE_ADAM
        elif self.optimizer == 'Adamax':
            rule = RULE_ADAMAX_METASGD
        else:
            raise NotImplementedError('This type of optimizer update operation is not yet implemented')
        updated = dict()
        for key in names_grads_wrt_params_dict.keys():
            g = names_grads_wrt_params_dict[key]
            if g is None:
                if rule == RULE_SGD:   # reference :329-330 multiplies by None
                    raise TypeError("unsupported operand type(s) for *: 'Parameter' and 'NoneType'")
                continue
            lr = self.names_learning_rates_dict[key.replace(".", "-")]
            updated[key] = _InnerUpdate.apply(names_weights_dict[key], lr, g, rule, 0, True, self.state[key],
                                              self.ops)
        return updated

    def update_sgd(self, names_weights_dict, names_grads_wrt_params_dict):
        saved, self.optimizer = self.optimizer, 'SGD'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict)
        finally:
            self.optimizer = saved

    def update_adam(self, names_weights_dict, names_grads_wrt_params_dict, amsgrad=False):
        saved, self.optimizer = self.optimizer, 'Adam'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict)
        finally:
            self.optimizer = saved

    def update_adamax(self, names_weights_dict, names_grads_wrt_params_dict):
        saved, self.optimizer = self.optimizer, 'Adamax'
        try:
            return self.update_params(names_weights_dict, names_grads_wrt_params_dict)
        finally:
            self.optimizer = saved
