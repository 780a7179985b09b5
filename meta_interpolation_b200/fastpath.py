"""Graph-captured executor of the MAML inner loop (the hot path proper).

For one task (reference meta_learning_system.py:366-461) it runs

    K x [ support step: both support triplets batched as N=2 ->
            forward, L1/MSE value+gradient, backward, and the inner update fused into the
            weight-gradient finishing kernel (w <- w - lr*g written straight into the
            fast-weight arena; no elementwise launch between inner steps) ]
    (+ per-step query passes when the MAML++ multi-step loss is active)
    1 x [ query pass: forward, loss, backward; outer gradients accumulated with scale
            1/B into a flat meta-gradient buffer ]

Each bracket is one CUDA graph, captured the second time it is needed and replayed
afterwards; frames are copied into static input buffers before a replay.  Fast
weights live in ONE arena updated in place after step 0 (step 0 reads the
meta-parameters and writes the arena), so three support graphs/two query graphs cover
any K.  Un-routed tensors read the meta arena and skip
their dead inner-loop weight gradients; the query pass differentiates everything.

Task-level concurrency (decision 3): the deep layers of a single task
launch only 24-96 CTAs on 148 SMs, and fast weights differ per task so tasks cannot be
batched into N.  Tasks are therefore dealt round-robin to ``task_streams`` LANES; a lane
owns a CUDA stream, its fast-weight arena, its graphs, its split-K workspace and its own
meta-gradient accumulators, so two tasks' graphs run concurrently with no shared mutable
state; the lanes' accumulators are summed once per meta-batch.

First-order outer gradients follow SURVEY Appendix E4:
  LSLR fixed lr      dL/dtheta = G
  LSLR learnable lr  + dL/dlr[t][k] = -<g_k[t], G[t]>
  Meta-SGD (SGD)     + dL/dalpha    = -(sum_k g_k) (.) G
  Meta-SGD (Adamax)  + dL/dalpha    = -((theta - w_K) / alpha) (.) G      (sum of the K update directions)
  multi-step loss    the above per step with weight w_k.

The Adam / Adamax inner rules (reference inner_loop_optimizers.py:150-244, :335-426) cannot be folded into the
weight-gradient epilogue (they need the whole gradient tensor's moments first), so their support graphs store the
step's gradients into a lane arena and apply ONE flat ``mi_inner_update`` launch over the whole fast-weight arena,
followed by the rotation of the routed filters; bias corrections and the lr column are baked per step, so these
rules capture one support graph per inner step.
"""
import gc
import os

import torch

from . import utils
from .arena import Arena
from .inner_loop_optimizers import RULE_ADAM, RULE_ADAMAX_LSLR, RULE_ADAMAX_METASGD, RULE_SGD
from .ops import WG_ACCUM, WG_SGD_SCALAR, WG_SGD_TENSOR, WG_STORE, WgradSpec
from .tape import ConvParam, Tape

LOSS_KIND = {'L1': 0, 'MSE': 1}


class _Sink:
    """Weight-gradient policy handed to the tape."""

    def __init__(self, lane, mode, step=0, scale=1.0):
        self.lane, self.mode, self.step, self.scale = lane, mode, step, scale

    def weight_grad(self, p, x, dy, k):
        lane = self.lane
        fp = lane.fp
        net = fp.net
        wn, bn = p.name + ".weight", p.name + ".bias"
        has_b = p.b is not None
        ldw = p.w.stride(2)
        if self.mode == 'inner':
            if not net.is_routed(wn):
                return                       # dead work in support passes (Q1/Q2/Q2b)
            if fp.rule != RULE_SGD:          # moment rules: store g, the flat update follows the backward pass
                garena = lane.gstep
                spec = WgradSpec(WG_STORE, grad_w=garena.kernel_view(wn),
                                 grad_b=garena.kernel_view(bn) if has_b else None)
                fp.ops.conv_wgrad(x, dy, k, ldw, spec)
                return
            fast = lane.fast
            spec = WgradSpec(w_in=p.w, b_in=p.b, w_out=fast.kernel_view(wn),
                             b_out=fast.kernel_view(bn) if has_b else None, wt_out=lane.wt_buffer(p.name),
                             wr_out=lane.fast_r.kernel_view(wn) if fp.ops.tf32_rn else None)
            if fp.metasgd:
                spec.mode = WG_SGD_TENSOR
                spec.lr_w = fp.sys.alpha.kernel_view(wn)
                spec.lr_b = fp.sys.alpha.kernel_view(bn) if has_b else None
                spec.gsum_w = lane.gsum.kernel_view(wn)
                spec.gsum_b = lane.gsum.kernel_view(bn) if has_b else None
            else:
                spec.mode = WG_SGD_SCALAR
                iw = net.layout.index(wn)
                spec.lr_w = lane.cur_lr[iw:iw + 1]
                if has_b:
                    ib = net.layout.index(bn)
                    spec.lr_b = lane.cur_lr[ib:ib + 1]
                if fp.learnable_lr:
                    garena = lane.gsteps[self.step]
                    spec.grad_w = garena.kernel_view(wn)
                    spec.grad_b = garena.kernel_view(bn) if has_b else None
            fp.ops.conv_wgrad(x, dy, k, ldw, spec)
        elif self.mode == 'accum':
            garena = lane.acc_theta
            spec = WgradSpec(WG_ACCUM, scale=self.scale, grad_w=garena.kernel_view(wn),
                             grad_b=garena.kernel_view(bn) if has_b else None)
            fp.ops.conv_wgrad(x, dy, k, ldw, spec)
        else:   # 'store': per-task query gradient G (needed for lr / alpha outer gradients)
            garena = lane.gquery
            spec = WgradSpec(WG_STORE, grad_w=garena.kernel_view(wn), grad_b=garena.kernel_view(bn) if has_b else None)
            fp.ops.conv_wgrad(x, dy, k, ldw, spec)

    def bn_targets(self, name):
        """(dgamma, dbeta, mode, scale) for a frozen batch norm, or None when its gradient is dead work
        (never routed from the fast weights: support passes skip it)."""
        lane = self.lane
        wn, bn = name + ".weight", name + ".bias"
        if self.mode == 'inner':
            assert not lane.fp.net.is_routed(wn)
            return None
        if self.mode == 'accum':
            g = lane.acc_theta
            return g.kernel_view(wn), g.kernel_view(bn), WG_ACCUM, self.scale
        g = lane.gquery
        return g.kernel_view(wn), g.kernel_view(bn), WG_STORE, 1.0


class _Program:
    """One capturable unit: static inputs, a body, static outputs."""

    def __init__(self, lane, body, n, h, w):
        dev = lane.fp.ops.device
        self.f0 = torch.zeros(n, 3, h, w, device=dev)
        self.f1 = torch.zeros(n, 3, h, w, device=dev)
        self.tgt = torch.zeros(n, 3, h, w, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.terms = torch.zeros(max(1, len(lane.fp.loss_terms)), device=dev)   # per-term values (multi-term losses)
        self.pred = None
        self.body = body
        self.graph = None
        self.calls = 0
        self.kernels = 0
        self.lane = lane

    def run(self):
        fp = self.lane.fp
        ops = fp.ops
        self.calls += 1
        if not fp.use_graphs or self.calls == 1:
            self.body(self)                      # eager (also sizes workspaces before a capture)
            return
        if self.graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = int(ops.lib.mi_launch_count())
            # no cyclic-GC pass while the stream is capturing: collecting an unreachable system of an earlier
            # meta-batch would destroy its CUDA graphs / free their pools in the middle of this capture
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, stream=self.lane.capture_stream):
                    self.body(self)
            finally:
                if gc_was_on:
                    gc.enable()
            self.kernels = int(ops.lib.mi_launch_count()) - n0   # recorded, not executed, during capture
            ops.replayed_launches -= self.kernels
            self.graph = g
        self.graph.replay()
        ops.replayed_launches += self.kernels


class _Lane:
    """Everything one in-flight task mutates: stream, fast weights, graphs, workspace slot, accumulators."""

    def __init__(self, fp, index):
        self.fp = fp
        self.index = index
        ops, net, sysm = fp.ops, fp.net, fp.sys
        dev = ops.device
        lay = net.layout
        cuda = ops.name == 'cuda'
        self.stream = torch.cuda.Stream(device=dev) if (cuda and fp.n_lanes > 1) else None
        self.capture_stream = torch.cuda.Stream(device=dev) if cuda else None
        self.fast = Arena(lay, dev)
        # TF32-rounded shadow of the fast weights: what fprop reads (the exact copy above is what updates read/write)
        self.fast_r = Arena(lay, dev) if ops.tf32_rn else self.fast
        self.cur_lr = torch.zeros(len(net.param_names), device=dev)
        self.gsum = Arena(lay, dev) if (fp.metasgd and fp.rule == RULE_SGD) else None
        # Adam / Adamax inner rules: gradient of the step in flight and the per-task moments
        self.gstep = Arena(lay, dev) if fp.rule != RULE_SGD else None
        self.exp_avg = Arena(lay, dev) if fp.rule in (RULE_ADAM, RULE_ADAMAX_LSLR) else None
        self.exp_avg_sq = Arena(lay, dev) if fp.rule == RULE_ADAM else None
        self.gquery = Arena(lay, dev) if (fp.metasgd or fp.learnable_lr or fp.l2f) else None
        self.gamma = None        # L2F attenuation of the task in flight
        self.l2f_records = []    # (task embedding, dL/dgamma) per adapted task, consumed once per meta-batch
        self.gsteps = []
        self.dots = torch.zeros(len(net.param_names), device=dev)
        # what the reference's loss AverageMeters see (meta_learning_system.py:323-332): every query pass's loss dict
        self.log_sum = torch.zeros(1 + len(fp.loss_terms), device=dev)
        self.log_n = 0
        self.programs = {}
        self.fast_wt = {}    # rotated (dgrad-layout) copies of the fast weights, written by the fused update
        # lane 0 accumulates straight into the optimizer's flat gradient buffers; the others into private
        # buffers that are added once per meta-batch
        if index == 0:
            self.acc_theta = sysm.net_grad
            self.acc_alpha = sysm.alpha_grad if fp.metasgd else None
            self.acc_lr = sysm.lr_table_grad if fp.learnable_lr else None
        else:
            self.acc_theta = Arena(lay, dev)
            self.acc_alpha = Arena(lay, dev) if fp.metasgd else None
            self.acc_lr = torch.zeros_like(sysm.lr_table_grad) if fp.learnable_lr else None

    def wt_buffer(self, name):
        """Persistent dgrad-layout buffer of the routed conv ``name``: the weight-gradient finishing kernel of inner
        step k writes the rotated updated weight here, inner step k+1 (and the query pass) read it."""
        buf = self.fast_wt.get(name)
        if buf is None:
            w = self.fast.kernel_view(name + ".weight")
            cout, k, _, cin = w.shape
            buf = self.fp.ops.empty_weight(cin, cout, k)
            self.fast_wt[name] = buf
        return buf

    def zero_accumulators(self):
        if self.index == 0:
            return
        ops = self.fp.ops
        ops.fill(self.acc_theta.flat, 0.0)
        if self.acc_alpha is not None:
            ops.fill(self.acc_alpha.flat, 0.0)
        if self.acc_lr is not None:
            self.acc_lr.zero_()

    def reduce_into_system(self):
        if self.index == 0:
            return
        ops, sysm = self.fp.ops, self.fp.sys
        ops.axpby(self.acc_theta.flat, 1.0, sysm.net_grad.flat, 1.0)
        if self.acc_alpha is not None:
            ops.axpby(self.acc_alpha.flat, 1.0, sysm.alpha_grad.flat, 1.0)
        if self.acc_lr is not None:
            sysm.lr_table_grad.add_(self.acc_lr)


class FastPath:
    @staticmethod
    def supports(system):
        a = system.args
        if a.second_order and system.current_epoch > a.first_order_to_second_order_epoch:
            return False
        if a.optimizer not in ('SGD', 'Adam', 'Adamax'):
            return False
        if a.optimizer != 'SGD':
            # moment rules are graph-captured with an LSLR table (fixed or learnable), as Meta-SGD + Adamax (the
            # authors' scripts/run_sepconv.sh operating point) and as Meta-SGD + Adam with ONE inner step (the point of
            # scripts/run_voxelflow.sh and run_cain.sh; the reference itself fails for K >= 2)
            if a.attenuate:
                return False
            if a.metasgd and a.optimizer == 'Adam' and a.number_of_training_steps_per_iter != 1:
                return False
        if a.attenuate:
            # L2F (reference :231-272) is graph-captured for the SGD inner rule (LSLR fixed / learnable, Meta-SGD,
            # with or without the multi-step loss) when every tensor is in the inner-loop dict
            if len(system.get_inner_loop_parameter_dict(system.net.named_parameters())) != len(system.net.param_names):
                return False
        kinds = [t.split('*')[1] for t in a.loss.split('+')]
        if not all(k in LOSS_KIND or (k == 'Super' and a.model == 'superslomo') for k in kinds):
            return False
        if a.metasgd and a.optimizer == 'SGD' and any(not system.net.is_routed(n) for n in system.net.param_names):
            return False   # the reference itself fails here for K>=2
        return hasattr(system.net, 'build_graph')

    def __init__(self, system):
        self.sys = system
        self.net = system.net
        self.ops = system.ops
        a = system.args
        self.metasgd = bool(a.metasgd)
        self.learnable_lr = (not self.metasgd) and bool(a.learnable_per_layer_per_step_inner_loop_learning_rate)
        self.l2f = bool(a.attenuate)
        if a.optimizer == 'SGD':
            self.rule = RULE_SGD
        elif a.optimizer == 'Adam':
            self.rule = RULE_ADAM
        else:
            self.rule = RULE_ADAMAX_METASGD if self.metasgd else RULE_ADAMAX_LSLR
        self.K = a.number_of_training_steps_per_iter
        self.use_graphs = bool(system.use_cuda_graphs) and self.ops.name == 'cuda'
        dev = self.ops.device
        lay = self.net.layout
        self.seg = lay.segment_table().to(dev)
        self.routed_mask = torch.tensor([1.0 if self.net.is_routed(n) else 0.0 for n in self.net.param_names],
                                        device=dev)
        self.skip = torch.tensor([0 if self.net.is_routed(n) else 1 for n in self.net.param_names], dtype=torch.uint8,
                                 device=dev)
        self.numel = torch.tensor([float(lay.logical_numel(n)) for n in self.net.param_names], device=dev)
        terms = [(t.split('*')[1], float(t.split('*')[0])) for t in a.loss.split('+')]
        self.loss_terms = [(LOSS_KIND[k], w) for k, w in terms if k in LOSS_KIND]
        # the `Super` loss (loss.py:246-274) differentiates through SuperSloMo's auxiliary outputs on the tape
        self.super_weight = sum(w for k, w in terms if k == 'Super') if any(k == 'Super' for k, _ in terms) else None
        self.super_terms = system.criterion.super_terms
        self.term_names = [k for k, _ in terms]
        self.multi_term = len(terms) > 1
        self.meta_wt = {}
        self.meta_r = Arena(self.net.layout, self.ops.device) if self.ops.tf32_rn else self.net.arena
        # Lanes and the SM budget of a persistent tensor-core launch go together (measured on B200, SepConv 256x448
        # K=5, 8 tasks per step, profiles/r02_sm_budget_sweep.txt): 4 lanes x 148 CTAs 43.8 tasks/s, 4 x 74 47.1,
        # 8 x 50 49.0, 8 x 37 49.8, 8 x 24 50.4, 8 x 18 49.0, 8 x 12 42.7.  A conv CTA owns its SM (~220 KB of shared
        # memory), so launches of different lanes never share one; with a quarter of the SMs each, four of them run
        # side by side and every CTA walks four times as many tiles per set-up (barriers, TMEM, first operand loads,
        # epilogue drain), which is what the 15-25 us launches of this path are short of.
        lanes = os.environ.get('MI_B200_TASK_STREAMS', getattr(a, 'task_streams', 8))
        self.n_lanes = max(1, int(lanes)) if self.ops.name == 'cuda' else 1
        self.lanes = [_Lane(self, i) for i in range(self.n_lanes)]
        self.sm_budget = 0                       # CTAs per persistent launch for the current call (0 = every SM)

    def refresh_meta_wt(self):
        """dgrad needs the rotated/transposed filter; for meta-parameters (un-routed tensors in support passes,
        everything at step 0 and in K=0 query passes) it only changes with the outer step, so it is rebuilt once
        per meta-batch into persistent buffers instead of once per pass."""
        if self.ops.tf32_rn:        # one launch: the rounded shadow of the whole meta arena (what fprop reads)
            self.ops.round_tf32(self.net.arena.flat, out=self.meta_r.flat)
        for name in self.net.conv_names:
            w = self.net.arena.kernel_view(name + ".weight")
            buf = self.meta_wt.get(name)
            if buf is None:
                cout, k, _, cin = w.shape
                buf = self.ops.empty_weight(cin, cout, k)
                self.meta_wt[name] = buf
            self.ops.weight_to_dgrad(w, out=buf)

    # ------------------------------------------------------------------ graph bodies
    def _provider(self, lane, src):
        net, fast = self.net, lane.fast
        cache = {}

        def provider(name):
            p = cache.get(name)
            if p is None:
                if src == 'fast' and net.is_routed(name + ".weight"):
                    w = fast.kernel_view(name + ".weight")
                    b = fast.kernel_view(name + ".bias") if net._spec[name][4] else None
                    p = ConvParam(name, w, b)
                    p._wt = lane.wt_buffer(name)     # kept current by the fused update of the previous step
                    p._wr = lane.fast_r.kernel_view(name + ".weight")
                else:
                    p = net.meta_param(name)
                    p._wt = self.meta_wt.get(name)   # rotated copy refreshed once per meta-batch
                    p._wr = self.meta_r.kernel_view(name + ".weight")
                cache[name] = p
            return p

        return provider

    def _seed_loss(self, prog, tape, out, n_pairs, backward=True):
        """Loss value of the pass into ``prog.loss`` and, when ``backward``, the gradient seeds on the tape."""
        grad = self._loss(prog, out.data, n_pairs)
        if backward and grad is not None:
            out.grad = grad
        if self.super_weight is not None:
            self.super_terms.seed(tape, out, self.net.aux_vars, prog.tgt, prog.f0, prog.f1,
                                  self.super_weight * n_pairs, prog.loss, backward)

    def _loss(self, prog, pred, n_pairs):
        """Pixel-loss terms of the pass: value into ``prog.loss`` (and, when the loss string has several terms, each
        term's own value into ``prog.terms`` for the per-term log the reference keeps, loss.py:325-350)."""
        self.ops.fill(prog.loss, 0.0)
        if not self.loss_terms:
            return None
        grad = torch.empty_like(pred)
        if not self.multi_term:
            kind, weight = self.loss_terms[0]
            self.ops.loss_fwd_bwd(pred, prog.tgt, kind, weight * n_pairs, prog.loss, grad)
            return grad
        self.ops.fill(prog.terms, 0.0)
        for i, (kind, weight) in enumerate(self.loss_terms):
            g = grad if i == 0 else torch.empty_like(pred)
            self.ops.loss_fwd_bwd(pred, prog.tgt, kind, weight * n_pairs, prog.terms[i:i + 1], g)
            if i:
                self.ops.axpby(g, 1.0, grad, 1.0)
        prog.loss.add_(prog.terms.sum())
        return grad

    def _backward(self, tape):
        """Backward pass of a captured body: the per-layer finishing launches of the weight gradients (split-K reduction
        + fused inner-loop update) are collected and issued together after the last layer -- the dgrad of a layer runs
        before its weight gradient and nothing in the pass reads an updated weight, so the order is free."""
        self.ops.begin_deferred_wgrad()
        try:
            tape.backward()
        finally:
            self.ops.flush_deferred_wgrad()

    def _support_body(self, lane, src, step_slot):
        def body(prog):
            self.ops.set_workspace_slot(lane.index)
            sink = _Sink(lane, 'inner', step=step_slot)
            tape = Tape(self.ops, self._provider(lane, src), sink, vectors=self.net.meta_bn)
            out = self.net.build_graph(tape, prog.f0, prog.f1)
            self._seed_loss(prog, tape, out, prog.f0.shape[0])
            self._backward(tape)
            if self.rule != RULE_SGD:
                self._moment_update(lane, src, step_slot)
        return body

    def _moment_update(self, lane, src, step):
        """Adam / Adamax inner step ``step`` over the whole arena (un-routed tensors skipped: their gradient is
        None in the reference, inner_loop_optimizers.py:209 / :393), then the rotated copies of the new filters."""
        ops, net = self.ops, self.net
        w_in = net.arena.flat if src == 'meta' else lane.fast.flat
        if self.metasgd:
            lr, per_element, stride = self.sys.alpha.flat, True, 0
        else:
            lr, per_element, stride = self.sys.lr_table, False, self.sys.lr_table.shape[1]
        keep = self.learnable_lr
        if keep:
            lane.gsteps[step].flat.copy_(w_in)          # w_k, turned into w_k - w_{k+1} below
        ops.inner_update(w_in, lane.gstep.flat, lane.fast.flat,
                         lane.exp_avg.flat if lane.exp_avg is not None else None,
                         lane.exp_avg_sq.flat if lane.exp_avg_sq is not None else None,
                         lr, per_element, stride, 0 if self.metasgd else step, self.seg, self.skip, self.rule, step + 1)
        if keep:
            # learnable per-step lr: dL/dlr[t][k] = -<dir_k[t], G[t]> with dir_k = (w_k - w_{k+1}) / lr[t][k]
            ops.axpby(lane.fast.flat, -1.0, lane.gsteps[step].flat, 1.0)
        if ops.tf32_rn:
            ops.round_tf32(lane.fast.flat, out=lane.fast_r.flat)
        for name in net.conv_names:
            if net.is_routed(name + ".weight"):
                ops.weight_to_dgrad(lane.fast.kernel_view(name + ".weight"), out=lane.wt_buffer(name))

    def _embed_body(self, lane):
        """L2F task embedding (reference :231-255): both support triplets through the META weights, gradient of the
        summed support loss w.r.t. every tensor stored into the lane's gradient arena."""
        def body(prog):
            self.ops.set_workspace_slot(lane.index)
            sink = _Sink(lane, 'store')
            tape = Tape(self.ops, self._provider(lane, 'meta'), sink, vectors=self.net.meta_bn)
            out = self.net.build_graph(tape, prog.f0, prog.f1)
            self._seed_loss(prog, tape, out, prog.f0.shape[0])
            self._backward(tape)
        return body

    def _attenuate(self, lane, frames, task, h, w, support_idxs):
        """gamma = clamp(1 - gamma_mult * attenuator(emb), 0, 1); fast <- gamma (.) theta (reference :258-272)."""
        ops, sysm = self.ops, self.sys
        prog = self._program(lane, ('embed', h, w), self._embed_body(lane), len(support_idxs), h, w)
        for i, (a, b, c) in enumerate(support_idxs):
            prog.f0[i].copy_(frames[a][task])
            prog.f1[i].copy_(frames[c][task])
            prog.tgt[i].copy_(frames[b][task])
        prog.run()
        ops.fill(lane.dots, 0.0)
        ops.segment_dot(lane.gquery.flat, None, self.seg, lane.dots)
        emb = lane.dots / self.numel                       # per-tensor mean of the support gradient
        with torch.no_grad():
            gamma = (1 - sysm.gamma_mult * sysm.attenuator(emb)).clamp(0, 1).contiguous()
        lane.gamma = gamma
        ops.segment_scale(self.net.arena.flat, gamma, self.seg, None, lane.fast.flat, 1.0, False)
        if ops.tf32_rn:
            ops.round_tf32(lane.fast.flat, out=lane.fast_r.flat)
        for name in self.net.conv_names:                   # rotated copies of the attenuated weights for step 0
            if self.net.is_routed(name + ".weight"):
                ops.weight_to_dgrad(lane.fast.kernel_view(name + ".weight"), out=lane.wt_buffer(name))
        return emb

    def _l2f_outer(self, lane, emb, scale):
        """First-order outer gradients through the attenuation: dL/dtheta_i += gamma_i G_i on the adapted tensors
        (G_i itself elsewhere) and dL/dgamma_i = <G_i, theta_i>; the attenuator / gamma_mult gradients follow from
        dL/dgamma by autograd over the small MLP once per meta-batch."""
        ops = self.ops
        G = lane.gquery.flat
        ops.segment_scale(G, lane.gamma, self.seg, self.routed_mask, lane.acc_theta.flat, scale, True)
        ops.fill(lane.dots, 0.0)
        ops.segment_dot(G, self.net.arena.flat, self.seg, lane.dots)
        lane.l2f_records.append((emb.clone(), (lane.dots * self.routed_mask * scale).clone()))

    def _query_body(self, lane, src, mode, backward=True):
        def body(prog):
            self.ops.set_workspace_slot(lane.index)
            sink = _Sink(lane, mode, scale=1.0) if backward else None
            if mode == 'accum' and backward:
                sink.scale = prog.scale
            tape = Tape(self.ops, self._provider(lane, src), sink, vectors=self.net.meta_bn)
            out = self.net.build_graph(tape, prog.f0, prog.f1)
            prog.pred = out.data
            self._seed_loss(prog, tape, out, 1, backward)
            if backward:
                self._backward(tape)
        return body

    MAX_PROGRAMS_PER_LANE = 32   # ~5 programs serve one (frame size, SM share); see _program

    def _program(self, lane, key, body, n, h, w):
        """The captured program for `key` on this lane, least-recently-used bounded: every entry owns static input
        buffers and a CUDA graph with a private memory pool, and the key contains what is baked into the graph (frame
        size, the accumulation scale -- which changes with the multi-step-loss epoch and with a short last batch --
        and the SM share), so validation on many frame sizes or a long MSL run would otherwise grow without bound."""
        key = key + (self.sm_budget,)        # grid sizes are baked into the captured graph
        progs = lane.programs
        p = progs.pop(key, None)
        if p is None:
            if len(progs) >= self.MAX_PROGRAMS_PER_LANE:
                if self.ops.name == 'cuda':
                    torch.cuda.synchronize() # the evicted graph may still be executing on the lane's stream
                for old in list(progs)[:len(progs) - self.MAX_PROGRAMS_PER_LANE + 1]:
                    del progs[old]           # (dicts keep insertion order: the front is the least recently used)
            p = _Program(lane, body, n, h, w)
        progs[key] = p                       # (re)inserted at the back: most recently used
        return p

    # ------------------------------------------------------------------ one task
    def _support_step(self, lane, frames, task, step, h, w, support_idxs):
        src = 'meta' if (step == 0 and not self.l2f) else 'fast'     # L2F: step 0 starts from gamma (.) theta
        slot = step if (self.learnable_lr or self.rule != RULE_SGD) else 0
        prog = self._program(lane, ('support', src, slot, h, w), self._support_body(lane, src, slot),
                             len(support_idxs), h, w)
        for i, (a, b, c) in enumerate(support_idxs):
            prog.f0[i].copy_(frames[a][task])
            prog.f1[i].copy_(frames[c][task])
            prog.tgt[i].copy_(frames[b][task])
        if self.rule != RULE_SGD:
            if step == 0:      # moments live for the K steps of one task (initialize_state per task, Q5)
                for arena in (lane.exp_avg, lane.exp_avg_sq):
                    if arena is not None:
                        self.ops.fill(arena.flat, 0.0)
        elif not self.metasgd:
            lane.cur_lr.copy_(self.sys.lr_table[:, step])
        elif step == 0:
            self.ops.fill(lane.gsum.flat, 0.0)
        prog.run()

    def _query(self, lane, frames, task, src, h, w, mode, scale, backward=True):
        ti = self.sys.target_idxs
        key = ('query', src, mode, backward, h, w, scale if mode == 'accum' else 0)
        prog = self._program(lane, key, self._query_body(lane, src, mode, backward), 1, h, w)
        prog.scale = scale
        prog.f0[0].copy_(frames[ti[0]][task])
        prog.f1[0].copy_(frames[ti[2]][task])
        prog.tgt[0].copy_(frames[ti[1]][task])
        prog.run()
        lane.log_sum[0:1].add_(prog.loss)
        if self.multi_term and self.loss_terms:
            lane.log_sum[1:].add_(prog.terms)
        lane.log_n += 1
        return prog

    def _outer_extras(self, lane, scale, steps_done):
        """alpha / lr outer gradients from the stored per-task query gradient G (Appx E4)."""
        ops = self.ops
        G = lane.gquery.flat
        if not self.l2f:                                                      # (L2F: gamma (.) G, see _l2f_outer)
            ops.axpby(G, scale, lane.acc_theta.flat, 1.0)                     # dL/dtheta += scale * G
        if self.metasgd and self.rule != RULE_SGD:
            # alpha is constant over the K steps, so the update directions sum to (theta - w_K) / alpha
            alpha = self.sys.alpha.flat
            dirs = torch.where(alpha != 0, (self.net.arena.flat - lane.fast.flat) / alpha, torch.zeros_like(alpha))
            ops.addcmul(lane.acc_alpha.flat, -scale, dirs, G)
        elif self.metasgd:
            ops.addcmul(lane.acc_alpha.flat, -scale, lane.gsum.flat, G)    # dL/dalpha -= scale * gsum (.) G
        elif self.learnable_lr:
            for j in range(steps_done):
                ops.fill(lane.dots, 0.0)
                ops.segment_dot(lane.gsteps[j].flat, G, self.seg, lane.dots)
                # only routed tensors are adapted; un-routed rows keep a zero gradient
                if self.rule == RULE_SGD:        # gsteps[j] = g_j
                    lane.acc_lr[:, j].add_(lane.dots * self.routed_mask, alpha=-scale)
                else:                            # gsteps[j] = w_j - w_{j+1} = lr[:, j] * dir_j
                    lane.acc_lr[:, j].add_(lane.dots * self.routed_mask / self.sys.lr_table[:, j], alpha=-scale)

    def adapt_and_query(self, lane, frames, task, num_steps, epoch, training, scale, msl, msl_w):
        """Inner loop + query for one task on ``lane``.  Returns (task_loss tensor[1], pred [1,3,H,W])."""
        h, w = frames[0].shape[2], frames[0].shape[3]
        sysm = self.sys
        support_idxs = sysm.support_idxs
        if self.learnable_lr:      # (also when evaluating: support graphs are shared between training and evaluation)
            while len(lane.gsteps) < num_steps:
                lane.gsteps.append(Arena(self.net.layout, self.ops.device))
        extras = training and (self.metasgd or self.learnable_lr)
        task_loss = torch.zeros(1, device=self.ops.device)
        prog = None
        emb = self._attenuate(lane, frames, task, h, w, support_idxs) if self.l2f else None
        for step in range(num_steps):
            self._support_step(lane, frames, task, step, h, w, support_idxs)
            if msl:
                wk = float(msl_w[step])
                if extras or (self.l2f and training):
                    prog = self._query(lane, frames, task, 'fast', h, w, 'store', 1.0)
                    if self.l2f:
                        self._l2f_outer(lane, emb, scale * wk)
                    if extras:
                        self._outer_extras(lane, scale * wk, step + 1)
                else:
                    prog = self._query(lane, frames, task, 'fast', h, w, 'accum', scale * wk)
                task_loss += wk * prog.loss
        if not msl:
            src = 'fast' if (num_steps > 0 or self.l2f) else 'meta'
            if not training:
                prog = self._query(lane, frames, task, src, h, w, 'none', 0.0, backward=False)
            elif self.l2f:
                prog = self._query(lane, frames, task, src, h, w, 'store', 1.0)
                self._l2f_outer(lane, emb, scale)
                if extras:
                    self._outer_extras(lane, scale, num_steps)
            elif extras:
                prog = self._query(lane, frames, task, src, h, w, 'store', 1.0)
                self._outer_extras(lane, scale, num_steps)
            else:
                prog = self._query(lane, frames, task, src, h, w, 'accum', scale)
            task_loss += prog.loss
        return task_loss, prog.pred.clone()

    # ------------------------------------------------------------------ meta-batch drivers
    def _run_tasks(self, frames, task_ids, num_steps, epoch, training, scale, msl, msl_w):
        """Deal the tasks round-robin to the lanes; lanes run concurrently on their own streams."""
        multi = self.n_lanes > 1
        main = torch.cuda.current_stream(self.ops.device) if multi else None
        # host copy of the multi-step-loss weights, taken ONCE: reading a device scalar per step (`float(msl_w[step])`)
        # is a stream synchronisation on the lane being fed, i.e. the host could not enqueue work for the other lanes
        # and the tasks of an MSL run (BASELINE configs[4]) executed one after the other
        msl_w = [float(v) for v in (msl_w.tolist() if torch.is_tensor(msl_w) else msl_w)]
        for lane in self.lanes:          # (run_test_iter never reads the log: start every meta-batch from zero)
            if lane.log_n:
                lane.log_sum.zero_()
                lane.log_n = 0
        if multi:
            for lane in self.lanes:
                lane.stream.wait_stream(main)      # inputs, zeroed gradients and rotated weights are ready
        # lanes that will actually be busy share the SMs (see __init__): four ways with eight or more tasks in flight,
        # two ways below that -- with as many tasks as shares the end of the meta-batch, where lanes run dry one by one,
        # would leave most SMs idle (measured at 4 tasks: superslomo 35.3 / 37.8 / 37.1 and cain 18.5 / 19.7 / 18.0
        # tasks/s for 1 / 2 / 4 shares, profiles/r02_sm_budget_sweep.txt)
        busy = min(len(task_ids), self.n_lanes)
        ways = 4 if busy >= 8 else (2 if busy >= 2 else 1)
        ways = getattr(self, 'force_sm_shares', None) or ways     # (bench.py: instrument one task at the 8-task share)
        self.sm_budget = self.ops.set_sm_budget(0) // ways if ways > 1 else 0
        self.ops.set_sm_budget(self.sm_budget)
        try:
            return self._deal_tasks(frames, task_ids, num_steps, epoch, training, scale, msl, msl_w, multi, main)
        finally:
            self.ops.set_sm_budget(0)

    def _deal_tasks(self, frames, task_ids, num_steps, epoch, training, scale, msl, msl_w, multi, main):
        losses_dev, preds = [None] * len(task_ids), [None] * len(task_ids)
        for i, t in enumerate(task_ids):
            lane = self.lanes[i % self.n_lanes]
            if multi:
                with torch.cuda.stream(lane.stream):
                    losses_dev[i], preds[i] = self.adapt_and_query(lane, frames, t, num_steps, epoch, training, scale,
                                                                   msl, msl_w)
            else:
                losses_dev[i], preds[i] = self.adapt_and_query(lane, frames, t, num_steps, epoch, training, scale, msl,
                                                               msl_w)
        if multi:
            for lane in self.lanes:
                main.wait_stream(lane.stream)
        return losses_dev, preds

    def _logged_losses(self):
        """{'total': ..., '<term>': ...}: averages over every query pass of the meta-batch, as the reference's
        ``update_loss_metrics`` / ``get_across_task_loss_metrics`` report them (:323-344), then reset."""
        n = sum(lane.log_n for lane in self.lanes)
        out = {}
        if n:
            tot = (torch.stack([lane.log_sum for lane in self.lanes]).sum(0) / n).cpu().numpy()
            out['total'] = tot[0]
            pixel = [self.term_names[i] for i in range(len(self.term_names)) if self.term_names[i] in LOSS_KIND]
            if not self.multi_term:
                out[self.term_names[0]] = tot[0]
            else:
                for i, name in enumerate(pixel):
                    out[name] = tot[1 + i]
                if 'Super' in self.term_names:
                    out['Super'] = tot[0] - tot[1:].sum()
        for lane in self.lanes:
            lane.log_sum.zero_()
            lane.log_n = 0
        return out

    def _finish(self, frames, task_ids, losses_dev, preds, do_evaluation, msl_w, loss_name_terms):
        sysm = self.sys
        n_tasks = len(frames[0])
        metrics = {'psnr': utils.AverageMeter(), 'ssim': utils.AverageMeter()}
        per_task = [[] for _ in range(n_tasks)]
        ti = sysm.target_idxs
        for t, p in zip(task_ids, preds):
            per_task[t] = sysm._denorm(p)
            if do_evaluation:
                psnr, ssim = utils.calc_metrics(sysm._denorm(p).squeeze(0), sysm._denorm(frames[ti[1]][t]),
                                                ops=self.ops)
                metrics['psnr'].update(psnr)
                metrics['ssim'].update(ssim)
        losses = {'loss': torch.cat(losses_dev).mean() if losses_dev else torch.zeros((), device=self.ops.device)}
        losses.update(self._logged_losses())
        for idx, item in enumerate(msl_w):
            losses['loss_importance_vector_{}'.format(idx)] = item.detach().cpu().numpy()
        return losses, per_task, metrics

    def train_iter(self, frames, epoch, task_ids, n_global, do_evaluation):
        """``task_ids``: the tasks of the meta-batch this rank adapts (possibly none); ``n_global``: size of the whole
        meta-batch -- every outer gradient is scaled by 1/n_global so the ranks' SUM is the reference's mean."""
        sysm, a = self.sys, self.sys.args
        msl_w = sysm.get_per_step_loss_importance_vector()
        msl = a.use_multi_step_loss_optimization and epoch < a.multi_step_loss_num_epochs
        scale = 1.0 / n_global
        sysm.optimizer.zero_grad()
        for g in sysm._groups:
            g.dirty = True
        for lane in self.lanes:
            lane.zero_accumulators()
        self.refresh_meta_wt()
        losses_dev, preds = self._run_tasks(frames, task_ids, self.K, epoch, True, scale, msl, msl_w)
        for lane in self.lanes:
            lane.reduce_into_system()
        if self.l2f:
            # attenuator / gamma_mult gradients: gamma_t = f(emb_t) re-evaluated under autograd with the embedding
            # as a constant (the reference's embedding carries no graph either: create_graph=False, :246-247)
            for lane in self.lanes:
                for emb, dldg in lane.l2f_records:
                    gamma = (1 - sysm.gamma_mult * sysm.attenuator(emb)).clamp(0, 1)
                    gamma.backward(dldg)
                lane.l2f_records = []
            sysm.optimizer.gather_grads()
        names = [t.split('*')[1] for t in a.loss.split('+')]
        return self._finish(frames, task_ids, losses_dev, preds, do_evaluation, msl_w, names)

    def test_iter(self, frames):
        """Test-time adaptation of 4-frame clips (reference run_test_iter, :630-697): support triplets (0,1,2) and
        (1,2,3), query frames (1,2) with no ground truth; returns the raw predictions [1,3,H,W] per clip."""
        sysm, a = self.sys, self.sys.args
        msl_w = sysm.get_per_step_loss_importance_vector()
        task_ids = list(range(len(frames[0])))
        self.refresh_meta_wt()
        saved = (sysm.support_idxs, sysm.target_idxs)
        sysm.support_idxs, sysm.target_idxs = [[0, 1, 2], [1, 2, 3]], [1, 1, 2]    # (the query "target" is a dummy)
        try:
            _, preds = self._run_tasks(frames, task_ids, a.number_of_evaluation_steps_per_iter, 0, False, 0.0, False,
                                       msl_w)
        finally:
            sysm.support_idxs, sysm.target_idxs = saved
        return preds

    def eval_iter(self, frames, epoch):
        sysm, a = self.sys, self.sys.args
        msl_w = sysm.get_per_step_loss_importance_vector()
        task_ids = list(range(len(frames[0])))
        self.refresh_meta_wt()
        losses_dev, preds = self._run_tasks(frames, task_ids, a.number_of_evaluation_steps_per_iter, epoch, False, 0.0,
                                            False, msl_w)
        names = [t.split('*')[1] for t in a.loss.split('+')]
        return self._finish(frames, task_ids, losses_dev, preds, True, msl_w, names)
