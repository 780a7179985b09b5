"""Oracle: the six inner-loop update rules (test infrastructure, not product).

Functional restatement of reference ``inner_loop_optimizers.py``; every rule
maps ``(weights, grads, lrs, state, num_step)`` to a NEW dict and reproduces
the reference's quirks:

* LSLR-SGD      :136-147  keys with ``None`` grad are dropped from the result
* LSLR-Adam     :150-198  moments persist across the K steps of one task
* LSLR-Adamax   :201-244  ``exp_avg`` persists, ``exp_inf`` is rebound to a local (:236) => always 0 on entry
* MetaSGD-SGD   :324-332  no ``None`` check (reference raises TypeError)
* MetaSGD-Adam  :335-382
* MetaSGD-Adamax:385-426  both moments rebound to locals (:409,:418) => stateless
"""
import math

import torch

BETA1, BETA2, EPS = 0.9, 0.99, 1e-8  # inner_loop_optimizers.py:91-94, 279-282


def lr_key(name):
    return name.replace(".", "-")  # inner_loop_optimizers.py:100, 290


def _lr(lrs, key, num_step, per_step):
    v = lrs[lr_key(key)]
    return v[num_step] if per_step else v


def update_sgd(weights, grads, lrs, state, num_step, per_step):
    out = {}
    for key, g in grads.items():
        if g is None:
            if per_step:
                continue                      # LSLR :141-142
            raise TypeError("unsupported operand type(s) for *: 'Parameter' and 'NoneType'")  # Meta-SGD :329-330
        out[key] = weights[key] - _lr(lrs, key, num_step, per_step) * g
    return out


def update_adam(weights, grads, lrs, state, num_step, per_step):
    out = {}
    for key, g in grads.items():
        if g is None:
            continue
        st = state.setdefault(key, {})
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(weights[key])
            st["exp_avg_sq"] = torch.zeros_like(weights[key])
        st["step"] += 1
        bc1 = 1 - BETA1 ** st["step"]
        bc2 = 1 - BETA2 ** st["step"]
        st["exp_avg"].mul_(BETA1).add_(g, alpha=1 - BETA1)
        st["exp_avg_sq"].mul_(BETA2).addcmul_(g, g, value=1 - BETA2)
        denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(EPS)
        step_size = _lr(lrs, key, num_step, per_step) / bc1
        out[key] = weights[key] - step_size * st["exp_avg"] / denom
    return out


def update_adamax(weights, grads, lrs, state, num_step, per_step):
    out = {}
    for key, g in grads.items():
        if g is None:
            continue
        st = state.setdefault(key, {})
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(weights[key])
            st["exp_inf"] = torch.zeros_like(weights[key])
        st["step"] += 1
        if per_step:
            st["exp_avg"].mul_(BETA1).add_(g, alpha=1 - BETA1)       # LSLR :229 (in place, persists)
            exp_avg = st["exp_avg"]
        else:
            exp_avg = (BETA1 * st["exp_avg"]).add(g, alpha=1 - BETA1)  # Meta-SGD :409 (local, state stays 0)
        # exp_inf.mul_(beta2) acts on the stored zeros; the max is rebound to a local (:232-236 / :412-418)
        exp_inf = torch.maximum(st["exp_inf"].mul_(BETA2), g.abs().add(EPS))
        bc = 1 - BETA1 ** st["step"]
        clr = _lr(lrs, key, num_step, per_step) / bc
        out[key] = weights[key] - clr * exp_avg / exp_inf
    return out


RULES = {"SGD": update_sgd, "Adam": update_adam, "Adamax": update_adamax}


def update_params(optimizer, metasgd, weights, grads, lrs, state, num_step):
    """Dispatch of inner_loop_optimizers.py:115-133 (LSLR) / :303-321 (Meta-SGD)."""
    if optimizer not in RULES:
        raise NotImplementedError("This type of optimizer update operation is not yet implemented")
    return RULES[optimizer](weights, grads, lrs, state, num_step, per_step=not metasgd)
