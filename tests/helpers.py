import argparse
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_args(**kw):
    """Namespace with the reference's flags (config.py:14-77 defaults that matter on the hot path)."""
    d = dict(model='sepconv', loss='1*L1', optimizer='SGD', inner_lr=1e-5, outer_lr=1e-5, batch_size=1, mode='train',
             resume=True, number_of_training_steps_per_iter=1, number_of_evaluation_steps_per_iter=1, metasgd=False,
             attenuate=False, learnable_per_layer_per_step_inner_loop_learning_rate=False,
             enable_inner_loop_optimizable_bn_params=False, second_order=False, first_order_to_second_order_epoch=-1,
             use_multi_step_loss_optimization=False, multi_step_loss_num_epochs=1, random_seed=12345, cuda=False,
             num_gpu=0, pretrained_model=None, weight_decay=1e-4, load_checkpoint=False, exp_name='exp', resume_exp=None)
    d.update(kw)
    return argparse.Namespace(**d)


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    if fx.get("frames") is None and fx.get("frames_u8") is not None:
        # full-size fixtures keep their structured frames as uint8 (oracle/make_golden.py::normalise_u8)
        from oracle.make_golden import normalise_u8
        fx["frames"] = torch.stack(normalise_u8(fx["frames_u8"], fx["args"]["model"]))
    return fx


def golden_names():
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt"))


def digest(t):
    t = t.detach().double().reshape(-1).cpu()
    return torch.tensor([t.sum(), t.abs().sum(), (t * t).sum()], dtype=torch.float64), t[:8].clone().float()


def oracle_from_fixture(fx):
    from oracle import backbones as bb, maml
    a = fx["args"]
    init = bb.seeded_params(a["model"], a["random_seed"])
    if fx.get("weight_gain") is not None:
        init = {k: (v * fx["weight_gain"] if v.dim() == 4 else v) for k, v in init.items()}
    if fx.get("init_transform") is not None:
        from oracle.make_golden import apply_init_transform
        init = {k: v.clone() for k, v in init.items()}
        apply_init_transform(init.items(), fx["init_transform"])
    return maml.OracleSystem(a["model"], init, optimizer=a["optimizer"],
                             metasgd=a["metasgd"], num_steps=a["number_of_training_steps_per_iter"],
                             inner_lr=a["inner_lr"], outer_lr=a["outer_lr"],
                             learnable_lr=a["learnable_per_layer_per_step_inner_loop_learning_rate"], loss=a["loss"],
                             attenuate=a["attenuate"], use_msl=a["use_multi_step_loss_optimization"],
                             msl_epochs=a["multi_step_loss_num_epochs"], attenuator_state=fx.get("attenuator_state"))


def system_from_fixture(fx, ops, **extra):
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    args = make_args(**fx["args"], **extra)
    args.cuda = ops.name == "cuda"
    system = SceneAdaptiveInterpolation(args, ops=ops)
    if fx.get("attenuator_state") is not None:
        system.attenuator.load_state_dict(fx["attenuator_state"])
    if fx.get("weight_gain") is not None:      # fixture generated from a rescaled seeded init (oracle/make_golden.py)
        with torch.no_grad():
            for p in system.net.parameters():
                if p.dim() == 4:
                    p.mul_(fx["weight_gain"])
    if fx.get("init_transform") is not None:
        from oracle.make_golden import apply_init_transform
        apply_init_transform(system.net.named_parameters(), fx["init_transform"])
    return system
