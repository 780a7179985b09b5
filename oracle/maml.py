"""Oracle: per-task inner loop + outer query pass (test infrastructure, not product).

Restates reference ``meta_learning_system.py``:
``forward`` :346-472, ``get_per_step_loss_importance_vector`` :186-210,
``get_task_embeddings`` :231-255, ``attenuate_init`` :258-272,
``apply_inner_loop_update`` :275-321, ``net_forward`` :475-509,
``meta_update`` :551-574, ``run_train_iter`` :584-606,
on plain ATen CPU ops with torch autograd, first-order only.
It is the checker for the CUDA path and the timed CPU baseline of ``bench.py``.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import backbones as bb
from . import inner_rules


def msl_weights(num_steps, epoch, msl_epochs):
    """meta_learning_system.py:186-210 (float64 numpy then cast, as the reference)."""
    if num_steps == 0:
        return torch.ones(1)
    w = np.ones(num_steps) * (1.0 / num_steps)
    decay = 1.0 / num_steps / msl_epochs
    floor = 0.03 / num_steps
    for i in range(num_steps - 1):
        w[i] = np.maximum(w[i] - epoch * decay, floor)
    w[-1] = np.minimum(w[-1] + epoch * (num_steps - 1) * decay, 1.0 - (num_steps - 1) * floor)
    return torch.Tensor(w)


def criterion(loss_spec, out, target, extras=None):
    """loss.py:325-350 for the L1/MSE/Super terms: {'L1': w*l, ..., 'total': sum}.  ``extras`` (Super only) =
    (aux dict of the superslomo forward, I0, I1, vgg16 conv4_3 state)."""
    losses = {}
    total = 0
    for term in loss_spec.split("+"):
        weight, kind = term.split("*")
        if kind == "L1":
            l = (out - target).abs().mean()
        elif kind == "MSE":
            l = ((out - target) ** 2).mean()
        elif kind == "Super":
            from .super_loss import super_loss
            aux, i0, i1, vgg_state = extras
            l = super_loss(out, target, aux, i0, i1, vgg_state)
        else:
            raise NotImplementedError(kind)
        losses[kind] = float(weight) * l
        total = total + losses[kind]
    losses["total"] = total
    return losses


def quantize(img):
    return img.mul(255).clamp(0, 255).round()  # utils.py:171-172 with rgb_range=1


def psnr(pred, gt):
    """utils.py:175-186 + :195-199."""
    diff = (quantize(pred) - quantize(gt)).div(255)
    mse = diff.pow(2).mean() + 1e-8
    return -10 * math.log10(float(mse))


class OracleSystem:
    """Functional twin of the reference's SceneAdaptiveInterpolation for one backbone."""

    def __init__(self, model, params, *, optimizer="SGD", metasgd=False, num_steps=1, inner_lr=1e-5,
                 outer_lr=1e-5, learnable_lr=False, loss="1*L1", attenuate=False, use_msl=False,
                 msl_epochs=1, attenuator_state=None, vgg_state=None, weight_decay=1e-4):
        self.model = model
        self.vgg_state = vgg_state      # torchvision vgg16 conv4_3 weights for the Super loss
        self.backbone = bb.BACKBONES[model]
        self.params = OrderedDict((k, nn.Parameter(v.detach().clone())) for k, v in params.items())
        self.optimizer_name = optimizer
        self.metasgd = metasgd
        self.num_steps = num_steps
        self.loss_spec = loss
        self.attenuate = attenuate
        self.use_msl = use_msl
        self.msl_epochs = msl_epochs
        self.lrs = OrderedDict()
        for k, v in self.params.items():
            if metasgd:   # inner_loop_optimizers.py:287-291
                self.lrs[inner_rules.lr_key(k)] = nn.Parameter(torch.ones_like(v) * inner_lr)
            else:         # :97-102
                self.lrs[inner_rules.lr_key(k)] = nn.Parameter(torch.ones(num_steps + 1) * inner_lr,
                                                               requires_grad=learnable_lr)
        self.attenuator = None
        if attenuate:     # meta_learning_system.py:107-117
            n = len(self.params)
            self.attenuator = nn.Sequential(nn.Linear(n, n), nn.ReLU(inplace=True), nn.Linear(n, n), nn.Sigmoid())
            if attenuator_state is not None:
                self.attenuator.load_state_dict(attenuator_state)
            self.gamma_mult = nn.Parameter(torch.zeros(1))
        # outer optimizer, meta_learning_system.py:132-143 (same flag as the inner rule, Q6)
        tp = self.trainable_parameters()
        if optimizer == "Adam" and model == "voxelflow":
            # :133-136: Adam over net.get_optim_policies() -- the backbone's tensors only, torch's default betas,
            # weight_decay; the groups' lr_mult / decay_mult keys are never read by torch.optim.Adam
            self.optimizer = torch.optim.Adam(list(self.params.values()), lr=outer_lr, weight_decay=weight_decay)
        elif optimizer == "Adam":
            self.optimizer = torch.optim.Adam(tp, lr=outer_lr, betas=(0.9, 0.99))
        elif optimizer == "Adamax":
            self.optimizer = torch.optim.Adamax(tp, lr=outer_lr, betas=(0.9, 0.999))
        else:
            self.optimizer = torch.optim.SGD(tp, lr=outer_lr)

    def trainable_parameters(self):
        # registration order of the reference module: net, inner_loop_optimizer, attenuator, gamma_mult
        out = [p for p in self.params.values()]
        out += [p for p in self.lrs.values() if p.requires_grad]
        if self.attenuate:
            out += list(self.attenuator.parameters()) + [self.gamma_mult]
        return out

    # ------------------------------------------------------------------ pieces
    def net_forward(self, f0, f1, target, fast):
        if "Super" in self.loss_spec:        # meta_learning_system.py:499-501: the plugin's extra outputs feed the loss
            from .backbones_flow import superslomo_forward
            out, aux = superslomo_forward(f0, f1, fast, self.params, full=True)
            return criterion(self.loss_spec, out, target, (aux, f0, f1, self.vgg_state)), out
        out = self.backbone["forward"](f0, f1, fast, self.params)
        return criterion(self.loss_spec, out, target), out

    def _support_loss(self, frames, task, fast, support_idxs):
        total = 0
        for a, b, c in support_idxs:
            l, _ = self.net_forward(frames[a][task:task + 1], frames[c][task:task + 1], frames[b][task:task + 1], fast)
            total = total + l["total"]
        return total

    def inner_update(self, loss, fast, state, step):
        grads = torch.autograd.grad(loss, list(fast.values()), create_graph=False, allow_unused=True)
        grads = dict(zip(fast.keys(), grads))
        new = inner_rules.update_params(self.optimizer_name, self.metasgd, fast, grads, self.lrs, state, step)
        return new, grads

    # ------------------------------------------------------------------ iteration
    def forward(self, frames, epoch, num_steps, training=True, support_idxs=((0, 2, 4), (2, 4, 6)),
                target_idx=(2, 3, 4), record=None):
        n_tasks = frames[0].shape[0]
        total_losses, preds, psnrs = [], [], []
        w = msl_weights(self.num_steps, epoch, self.msl_epochs)
        msl = self.use_msl and training and epoch < self.msl_epochs
        for task in range(n_tasks):
            fast = OrderedDict(self.params)
            state = {}
            task_losses = []
            if self.attenuate:
                sl = self._support_loss(frames, task, fast, support_idxs)
                g = torch.autograd.grad(sl, list(fast.values()), create_graph=False, allow_unused=True)
                emb = torch.stack([x.mean() for x in g])
                gamma = 1 - self.gamma_mult * self.attenuator(emb)
                gamma = gamma.clamp(0, 1)  # reference clamps in place (Q11); same values/grad
                fast = OrderedDict((k, gamma[i] * v) for i, (k, v) in enumerate(fast.items()))
            q = lambda fw: self.net_forward(frames[target_idx[0]][task:task + 1], frames[target_idx[2]][task:task + 1],
                                            frames[target_idx[1]][task:task + 1], fw)
            for step in range(num_steps):
                sl = self._support_loss(frames, task, fast, support_idxs)
                fast, grads = self.inner_update(sl, fast, state, step)
                if record is not None:
                    record.setdefault("support_loss", []).append(float(sl))
                    if task == 0 and step == 0:
                        record["grads_step0"] = {k: (None if v is None else v.detach().clone()) for k, v in grads.items()}
                if msl:
                    tl, out = q(fast)
                    task_losses.append(w[step] * tl["total"])
            if not training:
                with torch.no_grad():
                    tl, out = q(fast)
                task_losses.append(tl["total"])
            elif not msl:
                tl, out = q(fast)
                task_losses.append(tl["total"])
            if record is not None and task == 0:
                record["fast_final"] = {k: v.detach().clone() for k, v in fast.items()}
            preds.append(bb.denormalise(self.model, out.detach()[0]).unsqueeze(0))
            psnrs.append(psnr(bb.denormalise(self.model, out.detach()[0]),
                              bb.denormalise(self.model, frames[target_idx[1]][task])))
            total_losses.append(torch.sum(torch.stack(task_losses)))
        loss = torch.mean(torch.stack(total_losses))
        return loss, preds, psnrs

    def run_train_iter(self, frames, epoch=0, record=None, step_optimizer=True):
        """meta_learning_system.py:584-606.  Returns (loss, preds, psnrs, outer_grads)."""
        loss, preds, psnrs = self.forward(frames, epoch, self.num_steps, training=True, record=record)
        self.optimizer.zero_grad()
        loss.backward()
        grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in self.params.items()}
        lr_grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in self.lrs.items()}
        if step_optimizer:
            self.optimizer.step()
        self.optimizer.zero_grad()
        return loss.detach(), preds, psnrs, dict(theta=grads, lr=lr_grads)

    def run_validation_iter(self, frames, epoch=0, num_steps=None):
        """meta_learning_system.py:608-627 (first-order, query under no_grad)."""
        num_steps = self.num_steps if num_steps is None else num_steps
        loss, preds, psnrs = self.forward(frames, epoch, num_steps, training=False)
        return loss.detach(), preds, psnrs

    def run_test_iter(self, frames, num_steps=None):
        """meta_learning_system.py:630-697: 4-frame clips, support triplets (0,1,2) and (1,2,3), then the frame
        between frames 1 and 2 from the adapted weights.  Only superslomo is de-normalised (:686-690); every other
        backbone returns the raw network output [3,H,W]."""
        num_steps = self.num_steps if num_steps is None else num_steps
        support_idxs = ((0, 1, 2), (1, 2, 3))
        outs = []
        for task in range(frames[0].shape[0]):
            fast = OrderedDict(self.params)
            state = {}
            if self.attenuate:
                sl = self._support_loss(frames, task, fast, support_idxs)
                g = torch.autograd.grad(sl, list(fast.values()), create_graph=False, allow_unused=True)
                emb = torch.stack([x.mean() for x in g])
                gamma = (1 - self.gamma_mult * self.attenuator(emb)).clamp(0, 1)
                fast = OrderedDict((k, gamma[i] * v) for i, (k, v) in enumerate(fast.items()))
            for step in range(num_steps):
                sl = self._support_loss(frames, task, fast, support_idxs)
                fast, _ = self.inner_update(sl, fast, state, step)
            with torch.no_grad():
                out = self.backbone["forward"](frames[1][task:task + 1], frames[2][task:task + 1], fast, self.params)
            out = out.detach()[0]
            outs.append(bb.denormalise(self.model, out) if self.model == "superslomo" else out)
        return outs
