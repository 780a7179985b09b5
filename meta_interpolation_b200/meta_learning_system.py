"""Scene-adaptive interpolation meta-learner (drop-in for the reference's
``meta_learning_system.py:SceneAdaptiveInterpolation``).

Same constructor (an ``argparse.Namespace`` with the reference's flags), attributes
and methods as reference meta_learning_system.py:29-697; ``run_train_iter`` /
``run_validation_iter`` / ``run_test_iter`` return what the reference returns.

Two execution paths, selected per configuration:

* fast path (``fastpath.FastPath``): first-order MAML with an SGD inner rule (LSLR
  fixed/learnable or Meta-SGD), with or without the MAML++ multi-step loss.  Each
  inner step (2 support pairs batched as N=2: forward, loss, backward, fused
  ``w -= lr*g`` in the weight-gradient epilogue) and each query pass is one CUDA
  graph; per-task outer gradients are accumulated straight into the flat
  meta-gradient buffer (mathematically the reference's single ``loss.backward()``
  over B retained graphs, SURVEY section 7 decision 4).
* compat path: the reference's own control flow (dict of fast weights,
  ``torch.autograd.grad`` with ``allow_unused=True``, ``update_params``) running on
  the same kernels through ``torch.autograd.Function`` wrappers; used for the
  Adam/Adamax inner rules, L2F attenuation and by anything calling the pieces
  (``net_forward``, ``apply_inner_loop_update`` ...) individually.

Multi-GPU: one process per GPU; rank r adapts tasks [r*B/R, (r+1)*B/R) of the
meta-batch and the flat meta-gradient buffers are summed with one NCCL
all-reduce each before the (identical) outer step.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.distributed as dist

from . import utils
from .backbone import default_ops
from .inner_loop_optimizers import LSLRGradientDescentLearningRule, MetaSGDLearningRule
from .loss import Loss
from .outer_optim import FlatGroup, FusedOuterOptimizer


def set_torch_seed(seed):
    """reference meta_learning_system.py:16-26."""
    rng = np.random.RandomState(seed=seed)
    torch_seed = rng.randint(0, 999999)
    torch.manual_seed(seed=torch_seed)
    return rng


def _build_backbone(args, ops):
    resume = False if args.resume else True   # reference :52 (inverted flag kept)
    if args.model == 'sepconv':
        from .sepconv.model import MetaNetwork as MetaSepConv
        return MetaSepConv(resume=resume, strModel='l1', ops=ops)
    if args.model == 'cain':
        from .cain.model import MetaCAIN
        return MetaCAIN(depth=3, resume=resume, ops=ops)
    if args.model == 'rrin':
        from .rrin.model import MetaRRIN
        return MetaRRIN(level=3, resume=resume, ops=ops)
    if args.model == 'superslomo':
        from .superslomo.model import MetaSuperSloMo
        return MetaSuperSloMo(ops.device, resume=resume, ops=ops)
    if args.model == 'voxelflow':
        from .voxelflow.core.models.voxel_flow import MetaVoxelFlow
        return MetaVoxelFlow(args, resume=resume, ops=ops)
    # 'dain' needs its eight native extensions and is outside the hot path
    raise NotImplementedError('Model not implemented yet!')


class SceneAdaptiveInterpolation(nn.Module):
    def __init__(self, args, ops=None):
        super(SceneAdaptiveInterpolation, self).__init__()
        self.args = args
        self.ops = ops if ops is not None else default_ops()
        self.device = self.ops.device
        self.batch_size = args.batch_size
        self.use_cuda = args.cuda
        self.current_epoch = 0

        # frame indices of a septuplet: support triplets and the query triplet (reference :42-46)
        self.support_idxs = [[0, 2, 4], [2, 4, 6]]
        if args.mode == 'test':
            self.support_idxs = [[0, 1, 2], [1, 2, 3]]
        self.target_idxs = [2, 3, 4]

        self.rng = set_torch_seed(seed=args.random_seed)
        self.net = _build_backbone(args, self.ops)
        if args.model == 'superslomo':   # reference :70-73: outputs back to the 0~1 scale
            self._shift = torch.tensor([0.429, 0.431, 0.397], device=self.device).view(3, 1, 1)
            self.revNormalize = lambda img: img + self._shift
        elif args.model == 'voxelflow':  # reference :78-79
            self.mean = torch.FloatTensor([0.5 * 255] * 3).to(self.device).unsqueeze(1).unsqueeze(2)
            self.std = torch.FloatTensor([0.5 * 255] * 3).to(self.device).unsqueeze(1).unsqueeze(2)

        self.inner_learning_rate = args.inner_lr
        if self.args.metasgd:
            self.inner_loop_optimizer = MetaSGDLearningRule(device=self.device, optimizer=self.args.optimizer,
                                                            init_learning_rate=self.inner_learning_rate)
        else:
            self.inner_loop_optimizer = LSLRGradientDescentLearningRule(
                device=self.device, optimizer=self.args.optimizer, init_learning_rate=self.inner_learning_rate,
                total_num_inner_loop_steps=self.args.number_of_training_steps_per_iter,
                use_learnable_learning_rates=self.args.learnable_per_layer_per_step_inner_loop_learning_rate)
        self.inner_loop_optimizer._ops = self.ops

        names_weights_dict = self.get_inner_loop_parameter_dict(params=self.net.named_parameters())
        self.inner_loop_optimizer.initialize(names_weights_dict=names_weights_dict)

        if self.args.attenuate:   # L2F attenuator (reference :107-117)
            num_layers = len(names_weights_dict.keys())
            self.attenuator = nn.Sequential(
                nn.Linear(num_layers, num_layers), nn.ReLU(inplace=True),
                nn.Linear(num_layers, num_layers), nn.Sigmoid()).to(device=self.device)
            self.gamma_mult = nn.Parameter(torch.zeros(1, device=self.device))

        self._build_flat_groups()
        kind = self.args.optimizer if self.args.optimizer in ('Adam', 'Adamax') else 'SGD'
        betas = (0.9, 0.999) if kind == 'Adamax' else (0.9, 0.99)
        if kind == 'Adam' and args.model == 'voxelflow':
            # reference :133-136: Adam over net.get_optim_policies() with weight_decay and torch's default betas.
            # The policies hold the backbone's tensors only (conv weights / conv bias / BN scale+shift), so learnable
            # inner rates, Meta-SGD alphas and the attenuator are NOT stepped at this operating point (the one of the
            # authors' scripts/run_voxelflow.sh); 'lr_mult' / 'decay_mult' are carried as group keys that stock Adam
            # never reads: every group steps with the same lr and decay.
            policies = self.net.get_optim_policies()
            self.optimizer = FusedOuterOptimizer(self._groups[:1], self.ops, kind, lr=args.outer_lr,
                                                 betas=(0.9, 0.999), weight_decay=args.weight_decay)
            self.optimizer.set_reference_order([p for g in policies for p in g['params']], policies)
            self._groups = self._groups[:1]        # what is stepped is what is all-reduced
        else:
            self.optimizer = FusedOuterOptimizer(self._groups, self.ops, kind, lr=args.outer_lr, betas=betas)
            # checkpoints exchange optimizer state in the order of the reference's Adam(self.trainable_parameters())
            self.optimizer.set_reference_order(list(self.trainable_parameters()))
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer=self.optimizer, mode='min', factor=0.2,
                                                                    patience=5)
        self.criterion = Loss(args, ops=self.ops)
        self._load_weights()
        self._fast = None
        self.use_fast_path = getattr(args, 'fast_path', True)
        self.use_cuda_graphs = getattr(args, 'cuda_graphs', True)

    def _load_weights(self):
        """reference :154-170: ``--resume`` restores checkpoint/<exp>/checkpoint.pth (model_best.pth in val / test
        mode); ``--pretrained_model`` then overlays a backbone file with the lossy matching rules.  Harnesses that run
        from the seeded random init with no files on disk (tests, bench.py, smoke) set ``args.load_checkpoint=False``."""
        args = self.args
        if args.resume and getattr(args, 'load_checkpoint', True):
            print('Resume training')
            utils.load_checkpoint(args, self, None)
        if getattr(args, 'pretrained_model', None) is not None:
            print('Loading pretrained model: %s' % args.pretrained_model)
            checkpoint = torch.load(args.pretrained_model, map_location=self.device, weights_only=False)
            if args.model == 'superslomo' and 'state_dictFC' in checkpoint.keys():
                # the original SuperSloMo release keeps its two U-Nets in separate dicts
                merged = {'flowComp.' + k: v for k, v in checkpoint['state_dictFC'].items()}
                merged.update({'arbTimeFlowIntrp.' + k: v for k, v in checkpoint['state_dictAT'].items()})
                utils.lossy_load_state_dict(self.net, merged)
            elif args.model == 'superslomo':
                utils.lossy_load_state_dict(self, checkpoint['state_dict'])
            else:
                utils.lossy_load_state_dict(self.net, checkpoint['state_dict'])

    # ------------------------------------------------------------------ flat buffers
    def _build_flat_groups(self):
        net = self.net
        own = dict(net.named_parameters())
        g_net = FlatGroup('net', net.arena.flat, [])
        from .arena import Arena
        self.net_grad = Arena(net.layout, self.device, data=g_net.grad)
        for n in net.param_names:
            g_net.members.append((own[n], self.net_grad.reference_view(n)))
        groups = [g_net]
        lr_params = list(self.inner_loop_optimizer.names_learning_rates_dict.values())
        if self.args.metasgd:
            self.alpha = Arena(net.layout, self.device)
            g_lr = FlatGroup('alpha', self.alpha.flat, [])
            self.alpha_grad = Arena(net.layout, self.device, data=g_lr.grad)
            for n in net.param_names:
                p = self.inner_loop_optimizer.names_learning_rates_dict[n.replace('.', '-')]
                view = self.alpha.reference_view(n)
                view.copy_(p.data)
                p.data = view
                g_lr.members.append((p, self.alpha_grad.reference_view(n)))
            groups.append(g_lr)
        else:
            g_lr = FlatGroup.pack('lslr', lr_params, self.device)
            k1 = self.args.number_of_training_steps_per_iter + 1
            self.lr_table = g_lr.flat[:len(lr_params) * k1].view(len(lr_params), k1)
            self.lr_table_grad = g_lr.grad[:len(lr_params) * k1].view(len(lr_params), k1)
            if self.args.learnable_per_layer_per_step_inner_loop_learning_rate:
                groups.append(g_lr)
        if self.args.attenuate:
            groups.append(FlatGroup.pack('l2f', list(self.attenuator.parameters()) + [self.gamma_mult], self.device))
        self._groups = groups

    # ------------------------------------------------------------------ reference helpers
    def get_per_step_loss_importance_vector(self):
        """MSL importance weights (reference :186-210; closed form SURVEY Appx E6)."""
        k = self.args.number_of_training_steps_per_iter
        if k == 0:
            return torch.ones(1, device=self.device)
        w = np.ones(shape=(k)) * (1.0 / k)
        decay = 1.0 / k / self.args.multi_step_loss_num_epochs
        floor = 0.03 / k
        for i in range(k - 1):
            w[i] = np.maximum(w[i] - (self.current_epoch * decay), floor)
        w[-1] = np.minimum(w[-1] + (self.current_epoch * (k - 1) * decay), 1.0 - ((k - 1) * floor))
        return torch.Tensor(w).to(device=self.device)

    def get_inner_loop_parameter_dict(self, params):
        """reference :213-228."""
        out = dict()
        for name, param in params:
            if param.requires_grad:
                if self.args.enable_inner_loop_optimizable_bn_params or "norm_layer" not in name:
                    out[name] = param
        return out

    def get_task_embeddings(self, frames, task_id, names_weights_copy):
        """L2F task embedding = per-tensor mean of the support gradient (reference :231-255)."""
        support_loss = 0
        for ind in self.support_idxs:
            _loss, _ = self.net_forward(frame0=frames[ind[0]][task_id].unsqueeze(0),
                                        frame1=frames[ind[2]][task_id].unsqueeze(0),
                                        target=frames[ind[1]][task_id].unsqueeze(0),
                                        weights=names_weights_copy, backup_running_statistics=True, training=True,
                                        num_step=0)
            support_loss = support_loss + _loss['total']
        self.net.zero_grad(names_weights_copy)
        grads = torch.autograd.grad(support_loss, names_weights_copy.values(), create_graph=False, allow_unused=True)
        return torch.stack([g.mean() for g in grads])

    def attenuate_init(self, task_embeddings, names_weights_copy):
        """gamma = clamp(1 - gamma_mult * attenuator(emb), 0, 1); theta_i <- gamma_i * theta_i (reference :258-272)."""
        gamma = 1 - self.gamma_mult * self.attenuator(task_embeddings)
        gamma = gamma.clamp(0, 1)
        return {k: gamma[i] * v for i, (k, v) in enumerate(names_weights_copy.items())}

    def apply_inner_loop_update(self, loss, names_weights_copy, use_second_order, current_step_idx):
        """reference :275-321 (first-order: SURVEY F10)."""
        if use_second_order:
            raise NotImplementedError('second-order MAML cannot work for these backbones even in the reference '
                                      '(grid_sampler / raw-kernel backward, SURVEY F10)')
        self.net.zero_grad(params=names_weights_copy)
        grads = torch.autograd.grad(loss, names_weights_copy.values(), create_graph=False, allow_unused=True)
        names_grads_copy = dict(zip(names_weights_copy.keys(), grads))
        return self.inner_loop_optimizer.update_params(names_weights_dict=names_weights_copy,
                                                       names_grads_wrt_params_dict=names_grads_copy,
                                                       num_step=current_step_idx)

    def update_loss_metrics(self, task_losses, target_loss):
        for key, value in target_loss.items():
            if key not in task_losses:
                task_losses[key] = utils.AverageMeter()
            task_losses[key].update(value.detach())     # kept on device; read once per meta-batch

    def get_across_task_loss_metrics(self, total_losses, specific_losses):
        losses = dict()
        losses['loss'] = torch.mean(torch.stack(total_losses))
        for key, meter in specific_losses.items():
            avg = meter.avg
            losses[key] = avg.cpu().numpy() if torch.is_tensor(avg) else avg
        return losses

    def net_forward(self, frame0, frame1, target, weights, backup_running_statistics, training, num_step):
        """reference :475-509."""
        kwargs = {'backup_running_statistics': backup_running_statistics, 'num_step': num_step}
        if self.criterion.super_terms is not None:
            # the Super loss differentiates through the plugin's auxiliary outputs: forward and loss are one step
            return self.criterion.forward_with_plugin(self.net, frame0, frame1, target, weights)
        output = self.net.forward(frame0, frame1, params=weights, **kwargs)
        if self.args.model == 'superslomo':
            output[1]['I0'], output[1]['I1'] = frame0, frame1
            losses = self.criterion(output[0], target, **output[1])
            output = output[0]
        else:
            losses = self.criterion(output, target)
        return losses, output

    def trainable_parameters(self):
        for param in self.parameters():
            if param.requires_grad:
                yield param

    # ------------------------------------------------------------------ task sharding
    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _local_tasks(self, n_tasks):
        """Tasks of the meta-batch this rank adapts: the contiguous block [r*B/R, (r+1)*B/R) (floor division, so a
        meta-batch that does not divide by the world size -- the short last batch of an epoch -- is split raggedly
        and a rank may get none).  Gradients are scaled by 1/B (the global count) on every rank, so the SUM
        all-reduce is the reference's mean over the whole meta-batch whatever the split."""
        rank, world = self._world()
        return list(range(rank * n_tasks // world, (rank + 1) * n_tasks // world))

    def _allreduce_logged(self, losses, metrics, n_local, n_tasks):
        """one scalar all-reduce so every rank logs the meta-batch's loss / PSNR / SSIM, not its shard's."""
        rank, world = self._world()
        if world == 1:
            return
        # the same keys on every rank, whether or not it adapted a task: 'loss', 'total' and the loss string's terms
        keys = ['loss', 'total'] + [t.split('*')[1] for t in self.args.loss.split('+')]
        vals = [float(n_local) * float(torch.as_tensor(losses[k]).detach().reshape(-1)[0])
                if (n_local and k in losses) else 0.0 for k in keys]
        vals += [metrics['psnr'].sum, metrics['ssim'].sum, float(metrics['psnr'].count)]
        buf = torch.tensor(vals, dtype=torch.float64, device=self.device)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        out = buf.tolist()
        for i, k in enumerate(keys):
            mean = out[i] / max(n_tasks, 1)
            losses[k] = torch.tensor(mean, dtype=torch.float32, device=self.device) if k == 'loss' else \
                np.float32(mean)
        n_eval = out[-1]
        for name, total in (('psnr', out[-3]), ('ssim', out[-2])):
            m = metrics[name]
            m.sum, m.count = total, int(n_eval)
            m.avg = total / n_eval if n_eval else 0

    # ------------------------------------------------------------------ forward (compat control flow)
    def _denorm(self, pred):
        """Prediction / target back to [0,1] (reference :434-447).  Accepts [3,H,W] or [1,3,H,W]."""
        if self.args.model == 'superslomo':
            return pred + self._shift
        if self.args.model == 'voxelflow':
            return (pred * self.std + self.mean) / 255.0
        return pred

    def forward(self, data_batch, epoch, use_second_order, use_multi_step_loss_optimization, num_steps,
                training_phase, do_evaluation=False, task_ids=None):
        """reference :346-472 (one meta-batch, serial over tasks)."""
        frames = data_batch
        n_tasks = len(frames[0])
        task_ids = list(range(n_tasks)) if task_ids is None else task_ids
        total_losses = []
        loss_accumulator = {'total': utils.AverageMeter()}
        metrics = {'psnr': utils.AverageMeter(), 'ssim': utils.AverageMeter()}
        per_task_target_preds = [[] for _ in range(n_tasks)]
        self.net.zero_grad()
        msl = use_multi_step_loss_optimization and training_phase and epoch < self.args.multi_step_loss_num_epochs
        ti = self.target_idxs
        per_step_loss_importance_vectors = self.get_per_step_loss_importance_vector()

        for task_id in task_ids:
            task_losses = []
            names_weights_copy = self.get_inner_loop_parameter_dict(self.net.named_parameters())
            self.inner_loop_optimizer.initialize_state()
            if self.args.attenuate:
                emb = self.get_task_embeddings(frames, task_id, names_weights_copy)
                names_weights_copy = self.attenuate_init(task_embeddings=emb, names_weights_copy=names_weights_copy)

            def query(step_kw):
                return self.net_forward(frame0=frames[ti[0]][task_id].unsqueeze(0),
                                        frame1=frames[ti[2]][task_id].unsqueeze(0),
                                        target=frames[ti[1]][task_id].unsqueeze(0), weights=names_weights_copy,
                                        backup_running_statistics=False, training=True, num_step=step_kw)

            for num_step in range(num_steps):
                support_loss = 0
                for ind in self.support_idxs:
                    _loss, _ = self.net_forward(frame0=frames[ind[0]][task_id].unsqueeze(0),
                                                frame1=frames[ind[2]][task_id].unsqueeze(0),
                                                target=frames[ind[1]][task_id].unsqueeze(0),
                                                weights=names_weights_copy,
                                                backup_running_statistics=(num_step == 0), training=True,
                                                num_step=num_step)
                    support_loss = support_loss + _loss['total']
                names_weights_copy = self.apply_inner_loop_update(loss=support_loss,
                                                                  names_weights_copy=names_weights_copy,
                                                                  use_second_order=use_second_order,
                                                                  current_step_idx=num_step)
                if msl:
                    target_loss, target_preds = query(num_step)
                    task_losses.append(per_step_loss_importance_vectors[num_step] * target_loss['total'])
                    self.update_loss_metrics(loss_accumulator, target_loss)

            if not training_phase:
                with torch.no_grad():
                    target_loss, target_preds = query(num_steps)
                task_losses.append(target_loss['total'])
                self.update_loss_metrics(loss_accumulator, target_loss)
            elif not msl:
                target_loss, target_preds = query(num_steps)
                task_losses.append(target_loss['total'])
                self.update_loss_metrics(loss_accumulator, target_loss)

            per_task_target_preds[task_id] = self._denorm(target_preds.detach())
            if do_evaluation:
                psnr, ssim = utils.calc_metrics(self._denorm(target_preds.detach()).squeeze(0),
                                                self._denorm(frames[ti[1]][task_id]), ops=self.ops)
                metrics['psnr'].update(psnr)
                metrics['ssim'].update(ssim)
            total_losses.append(torch.sum(torch.stack(task_losses)))
            if not training_phase:
                self.net.restore_backup_stats()

        losses = self.get_across_task_loss_metrics(total_losses=total_losses, specific_losses=loss_accumulator)
        for idx, item in enumerate(per_step_loss_importance_vectors):
            losses['loss_importance_vector_{}'.format(idx)] = item.detach().cpu().numpy()
        return losses, per_task_target_preds, metrics

    def train_forward_prop(self, data_batch, epoch, do_evaluation=False, task_ids=None):
        return self.forward(data_batch=data_batch, epoch=epoch,
                            use_second_order=self.args.second_order and epoch > self.args.first_order_to_second_order_epoch,
                            use_multi_step_loss_optimization=self.args.use_multi_step_loss_optimization,
                            num_steps=self.args.number_of_training_steps_per_iter, training_phase=True,
                            do_evaluation=do_evaluation, task_ids=task_ids)

    def evaluation_forward_prop(self, data_batch, epoch, task_ids=None):
        return self.forward(data_batch=data_batch, epoch=epoch, use_second_order=False,
                            use_multi_step_loss_optimization=True,
                            num_steps=self.args.number_of_evaluation_steps_per_iter, training_phase=False,
                            do_evaluation=True, task_ids=task_ids)

    def meta_update(self, loss):
        """reference :551-574: zero_grad; backward; step (gradients land in the flat buffers)."""
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.gather_grads()
        self._allreduce_grads()
        self.optimizer.step()

    def _allreduce_grads(self):
        rank, world = self._world()
        if world == 1:
            return
        for g in self._groups:
            dist.all_reduce(g.grad, op=dist.ReduceOp.SUM)

    # ------------------------------------------------------------------ fast path
    def fast_path(self):
        if self._fast is None:
            from .fastpath import FastPath
            self._fast = FastPath(self)
        return self._fast

    def fast_path_supported(self):
        from .fastpath import FastPath
        return self.use_fast_path and FastPath.supports(self)

    # ------------------------------------------------------------------ iterations
    def run_train_iter(self, data_batch, epoch, do_evaluation=False):
        """reference :584-606."""
        epoch = int(epoch)
        if self.current_epoch != epoch:
            self.current_epoch = epoch
        if not self.training:
            self.train()
        data_batch = [frame.to(device=self.device) for frame in data_batch]
        n_tasks = len(data_batch[0])
        task_ids = self._local_tasks(n_tasks)

        if self.fast_path_supported():
            losses, preds, metrics = self.fast_path().train_iter(data_batch, epoch, task_ids, n_tasks, do_evaluation)
            self._allreduce_grads()
            self.optimizer.step()
        elif task_ids:
            losses, preds, metrics = self.train_forward_prop(data_batch=data_batch, epoch=epoch,
                                                             do_evaluation=do_evaluation, task_ids=task_ids)
            # mean over the local tasks -> this rank's share of the mean over the whole meta-batch
            self.meta_update(loss=losses['loss'] * (len(task_ids) / n_tasks))
        else:      # more ranks than tasks: nothing to adapt here, but the collective and the step still happen
            losses, preds, metrics = self._empty_results(n_tasks)
            self.optimizer.zero_grad()
            self._allreduce_grads()
            self.optimizer.step()
        self._allreduce_logged(losses, metrics, len(task_ids), n_tasks)
        self.optimizer.zero_grad()
        self.zero_grad()
        return losses, preds, metrics

    def _empty_results(self, n_tasks):
        losses = {'loss': torch.zeros((), device=self.device), 'total': np.float32(0.0)}
        for idx, item in enumerate(self.get_per_step_loss_importance_vector()):
            losses['loss_importance_vector_{}'.format(idx)] = item.detach().cpu().numpy()
        return losses, [[] for _ in range(n_tasks)], {'psnr': utils.AverageMeter(), 'ssim': utils.AverageMeter()}

    def run_validation_iter(self, data_batch):
        """reference :608-627."""
        data_batch = [frame.to(device=self.device) for frame in data_batch]
        if self.fast_path_supported():
            return self.fast_path().eval_iter(data_batch, self.current_epoch)
        return self.evaluation_forward_prop(data_batch=data_batch, epoch=self.current_epoch)

    # ------------------------------------------------------------------ large frames
    @staticmethod
    def _halves(frames):
        h, w = frames[0].shape[-2:]
        if h > w:
            return [im[..., :h // 2, :] for im in frames], [im[..., h // 2:, :] for im in frames], -2
        return [im[..., :w // 2] for im in frames], [im[..., w // 2:] for im in frames], -1

    def _needs_tiling(self, frames):
        h, w = frames[0].shape[-2:]
        return h * w > 5e5 or (self.args.model == 'rrin' and h * w > 3e5)

    def run_validation_iter_tiled(self, data_batch):
        """``ExperimentBuilder.evaluation_iteration``'s frame splitting (experiment_builder.py:101-128): frames above
        5e5 pixels (3e5 for rrin) are halved along their longer side, recursively, each half adapted and predicted on
        its own, predictions concatenated and losses averaged.  Returns (losses, preds) like the reference's
        ``_eval_iter`` (its metrics are recomputed by the caller on the stitched frame)."""
        frames = [f.to(device=self.device) for f in data_batch]
        if not self._needs_tiling(frames):
            losses, outputs, _ = self.run_validation_iter(frames)
            losses['loss'] = losses['loss'].detach()
            return losses, outputs
        f0, f1, dim = self._halves(frames)
        l0, o0 = self.run_validation_iter_tiled(f0)
        l1, o1 = self.run_validation_iter_tiled(f1)
        outputs = [torch.cat([a, b], dim=dim) for a, b in zip(o0, o1)]
        losses = l0
        for k, v in l1.items():
            losses[k] = (v + losses[k]) / 2
        losses['loss'] = losses['loss'].detach()
        return losses, outputs

    def run_test_iter_tiled(self, data_batch):
        """``ExperimentBuilder.test_iteration`` (experiment_builder.py:153-172): ONE level of halving above 5e5 px."""
        frames = [f.to(device=self.device) for f in data_batch]
        h, w = frames[0].shape[-2:]
        if h * w <= 5e5:
            return self.run_test_iter(frames)
        f0, f1, dim = self._halves(frames)
        o0, o1 = self.run_test_iter(f0), self.run_test_iter(f1)
        return [torch.cat([a, b], dim=dim) for a, b in zip(o0, o1)]

    def run_test_iter(self, data_batch):
        """reference :630-697: 4-frame clips, support [[0,1,2],[1,2,3]], query (1,2) -> list of [3,H,W]."""
        if self.training:
            self.eval()
        frames = [frame.to(device=self.device) for frame in data_batch]
        if self.fast_path_supported():
            outs = self.fast_path().test_iter(frames)
            # reference :686-690 de-normalises superslomo only
            return [(self._denorm(o) if self.args.model == 'superslomo' else o).squeeze(0) for o in outs]
        preds = [[] for _ in range(len(frames[0]))]
        self.net.zero_grad()
        support_idxs = [[0, 1, 2], [1, 2, 3]]
        for task_id in range(len(frames[0])):
            names_weights_copy = self.get_inner_loop_parameter_dict(self.net.named_parameters())
            self.inner_loop_optimizer.initialize_state()
            saved, self.support_idxs = self.support_idxs, support_idxs
            try:
                if self.args.attenuate:
                    emb = self.get_task_embeddings(frames, task_id, names_weights_copy)
                    names_weights_copy = self.attenuate_init(task_embeddings=emb,
                                                             names_weights_copy=names_weights_copy)
            finally:
                self.support_idxs = saved
            num_step = 0
            for num_step in range(self.args.number_of_evaluation_steps_per_iter):
                support_loss = 0
                for ind in support_idxs:
                    _loss, _ = self.net_forward(frame0=frames[ind[0]][task_id].unsqueeze(0),
                                                frame1=frames[ind[2]][task_id].unsqueeze(0),
                                                target=frames[ind[1]][task_id].unsqueeze(0),
                                                weights=names_weights_copy,
                                                backup_running_statistics=(num_step == 0), training=True,
                                                num_step=num_step)
                    support_loss = support_loss + _loss['total']
                names_weights_copy = self.apply_inner_loop_update(loss=support_loss,
                                                                  names_weights_copy=names_weights_copy,
                                                                  use_second_order=self.args.second_order,
                                                                  current_step_idx=num_step)
            with torch.no_grad():
                output = self.net.forward(frames[1][task_id].unsqueeze(0), frames[2][task_id].unsqueeze(0),
                                          params=names_weights_copy, backup_running_statistics=False, training=True,
                                          num_step=num_step)
            if self.args.model == 'superslomo':
                output = output[0]
            output = output.detach()
            if self.args.model == 'superslomo':   # reference :686-690 de-normalises superslomo only
                output = self._denorm(output)
            preds[task_id] = output.squeeze(0)
            self.net.restore_backup_stats()
        return preds
