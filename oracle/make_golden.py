"""Generate tests/golden/*.pt by running the UNMODIFIED reference on CPU (container only).

    python -m oracle.make_golden            # needs /root/reference

For every case below the reference's ``SceneAdaptiveInterpolation`` (under the shims
of ``oracle/reference_shims.py``) runs one ``run_train_iter`` on seeded synthetic
frames; the oracle (``oracle/maml.py``) runs the same case and must agree with the
reference before anything is written (that is the pin of the oracle).  The fixture
stores the inputs, the reference's outputs and compact per-tensor digests of the
meta-gradients / post-step parameters (full tensors would be ~90 MB per case).
"""
import argparse
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (reference args overrides, frame size, batch)
    "sepconv_lslr_sgd_k2": (dict(model="sepconv", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=2), 64, 1),
    "sepconv_lslr_sgd_k1_b2_mse": (dict(model="sepconv", loss="1*MSE", optimizer="SGD",
                                        number_of_training_steps_per_iter=1), 48, 2),
    "sepconv_lslr_learnable_msl_k2": (dict(model="sepconv", loss="1*L1", optimizer="SGD",
                                           number_of_training_steps_per_iter=2,
                                           learnable_per_layer_per_step_inner_loop_learning_rate=True,
                                           use_multi_step_loss_optimization=True), 48, 1),
    "sepconv_lslr_adam_k2": (dict(model="sepconv", loss="1*L1", optimizer="Adam",
                                  number_of_training_steps_per_iter=2), 48, 1),
    "sepconv_metasgd_adamax_k2": (dict(model="sepconv", loss="1*L1", optimizer="Adamax", metasgd=True,
                                       number_of_training_steps_per_iter=2), 48, 1),
    "sepconv_l2f_sgd_k1": (dict(model="sepconv", loss="1*L1", optimizer="SGD", attenuate=True,
                                number_of_training_steps_per_iter=1), 48, 1),
    # BASELINE configs[0]: voxelflow, batch 1, 128x128, 1 inner step, CPU-runnable
    "voxelflow_lslr_sgd_k1_mse": (dict(model="voxelflow", loss="1*MSE", optimizer="SGD",
                                       number_of_training_steps_per_iter=1), 128, 1),
    # the operating point of the authors' scripts/run_voxelflow.sh: Meta-SGD with the Adam inner rule, K=1, and the
    # outer Adam built from net.get_optim_policies() with weight decay (meta_learning_system.py:133-136)
    "voxelflow_metasgd_adam_k1": (dict(model="voxelflow", loss="1*MSE", optimizer="Adam", metasgd=True,
                                       number_of_training_steps_per_iter=1), 64, 2),
    "voxelflow_lslr_sgd_k2_ragged": (dict(model="voxelflow", loss="1*L1", optimizer="SGD",
                                          number_of_training_steps_per_iter=2), (72, 88), 1),
    # configs[2] in miniature: superslomo Meta-SGD (SGD rule), K=2
    "superslomo_metasgd_sgd_k2": (dict(model="superslomo", loss="1*L1", optimizer="SGD", metasgd=True,
                                       number_of_training_steps_per_iter=2), 64, 2),
    "superslomo_lslr_sgd_k1_ragged": (dict(model="superslomo", loss="1*L1", optimizer="SGD",
                                           number_of_training_steps_per_iter=1), (72, 80), 1),
    # the loss of the authors' scripts/run_superslomo.sh (loss.py:246-274); the ImageNet VGG16 is not available
    # offline, so torchvision's vgg16() is handed a seeded random conv4_3 (oracle/super_loss.py) on both sides
    "superslomo_super_sgd_k1": (dict(model="superslomo", loss="1*Super", optimizer="SGD",
                                     number_of_training_steps_per_iter=1, _vgg_seed=0), 64, 1),
    # configs[4] in miniature: rrin MAML++ (multi-step loss + learnable per-step lr), K=2
    "rrin_msl_learnable_k2": (dict(model="rrin", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=2,
                                   learnable_per_layer_per_step_inner_loop_learning_rate=True,
                                   use_multi_step_loss_optimization=True), (64, 128), 1),
    "rrin_lslr_sgd_k1_ragged": (dict(model="rrin", loss="1*L1", optimizer="SGD",
                                     number_of_training_steps_per_iter=1), (72, 136), 1),
    # configs[3] in miniature: cain L2F (--attenuate)
    "cain_l2f_sgd_k1": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True,
                             number_of_training_steps_per_iter=1), 128, 1),
    "cain_lslr_sgd_k2_ragged": (dict(model="cain", loss="1*L1", optimizer="SGD",
                                     number_of_training_steps_per_iter=2), (120, 136), 1),
    # cain's 125 stacked xavier convs explode at the default init (|pred| ~ 1e2, loss ~ 20: SURVEY 8d), which amplifies
    # any rounding difference a hundredfold; the same case from a well-conditioned start (every 4-D weight of the
    # seeded init scaled by WEIGHT_GAIN *before* the reference runs; no reference code is touched)
    "cain_lslr_sgd_k2_gain04": (dict(model="cain", loss="1*L1", optimizer="SGD",
                                     number_of_training_steps_per_iter=2, _weight_gain=0.4), (120, 136), 1),
    "cain_l2f_sgd_k1_gain04": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True,
                                    number_of_training_steps_per_iter=1, _weight_gain=0.4), 128, 1),
}


# BASELINE.json configs[1..4] at their REAL frame size and inner-step count, one task each, on structured 8-bit
# frames (bench.synthetic_septuplets quantised to uint8 like the dataset's PNGs; stored as uint8, 2.4 MB per task).
# The reference needs 1-4 minutes of CPU per case; these are the pins of the configuration bench.py measures.
FULL_CASES = {
    "full_sepconv_c2_k5": (dict(model="sepconv", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=5),
                           ("u8", 256, 448, 100), 1),
    # the same configuration from an init that interpolates: the four filter Subnets' last biases hold a centred delta
    # of sqrt(0.5), so the prediction is ~ (frame0 + frame1) / 2 and PSNR sits near 30 dB, where |dPSNR| < 0.01 dB
    # is a hundred times tighter a bar than at the 5 dB of the raw seeded init
    "full_sepconv_c2_k5_delta": (dict(model="sepconv", loss="1*L1", optimizer="SGD",
                                      number_of_training_steps_per_iter=5, _init="sepconv_delta"),
                                 ("u8", 256, 448, 104), 1),
    "full_superslomo_c3_metasgd_k5": (dict(model="superslomo", loss="1*L1", optimizer="SGD", metasgd=True,
                                           number_of_training_steps_per_iter=5), ("u8", 256, 448, 101), 1),
    "full_rrin_c5_msl_k5": (dict(model="rrin", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=5,
                                 learnable_per_layer_per_step_inner_loop_learning_rate=True,
                                 use_multi_step_loss_optimization=True), ("u8", 256, 448, 102), 1),
    "full_cain_c4_l2f_k3_gain04": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True,
                                        number_of_training_steps_per_iter=3, _weight_gain=0.4),
                                   ("u8", 512, 512, 103), 1),
}


# run_test_iter (meta_learning_system.py:630-697): 4-frame clips, test-time adaptation, prediction between frames 1, 2
TEST_ITER_CASES = {
    "test_iter_sepconv_k2": (dict(model="sepconv", loss="1*L1", optimizer="SGD",
                                  number_of_evaluation_steps_per_iter=2), (48, 56), 2),
    "test_iter_superslomo_k2": (dict(model="superslomo", loss="1*L1", optimizer="SGD",
                                     number_of_evaluation_steps_per_iter=2), (64, 64), 2),
    "test_iter_voxelflow_k1": (dict(model="voxelflow", loss="1*MSE", optimizer="SGD",
                                    number_of_evaluation_steps_per_iter=1), (64, 64), 1),
    "test_iter_rrin_l2f_k1": (dict(model="rrin", loss="1*L1", optimizer="SGD", attenuate=True, mode="test",
                                   number_of_evaluation_steps_per_iter=1), (64, 72), 1),
}


def run_test_iter_case(name, over, size, batch):
    """The reference's run_test_iter on seeded 4-frame clips; the oracle must agree before the fixture is written."""
    from oracle import reference_shims as rs
    from oracle import maml
    system, args = rs.build_system(batch_size=batch, **over)
    frames = synthetic_frames(3, batch, size, over.get("model", "sepconv"))[:4]
    init = {k: v.detach().clone() for k, v in system.net.named_parameters()}
    att_state = {k: v.detach().clone() for k, v in system.attenuator.state_dict().items()} if args.attenuate else None
    if args.attenuate:          # gamma_mult starts at 0 (gamma == 1): move it so the attenuation is exercised
        with torch.no_grad():
            system.gamma_mult.fill_(0.5)
    ora = maml.OracleSystem(args.model, init, optimizer=args.optimizer, metasgd=args.metasgd,
                            num_steps=args.number_of_evaluation_steps_per_iter, inner_lr=args.inner_lr,
                            outer_lr=args.outer_lr, loss=args.loss, attenuate=args.attenuate,
                            attenuator_state=att_state)
    if args.attenuate:
        with torch.no_grad():
            ora.gamma_mult.fill_(0.5)
    mine = ora.run_test_iter(frames)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = system.run_test_iter(frames)
    pin = max((a - b).abs().max().item() for a, b in zip(mine, ref))
    assert pin <= 1e-6, (name, pin)
    fixture = dict(
        name=name, args={k: getattr(args, k) for k in (
            "model", "loss", "optimizer", "metasgd", "attenuate", "inner_lr", "outer_lr", "batch_size", "random_seed",
            "number_of_training_steps_per_iter", "number_of_evaluation_steps_per_iter",
            "learnable_per_layer_per_step_inner_loop_learning_rate", "use_multi_step_loss_optimization",
            "multi_step_loss_num_epochs", "second_order", "first_order_to_second_order_epoch",
            "enable_inner_loop_optimizable_bn_params")},
        frames=torch.stack(frames), outputs=torch.stack([r.detach() for r in ref]),
        oracle_vs_reference_maxabs=pin, attenuator_state=att_state, gamma_mult=0.5 if args.attenuate else None)
    path = os.path.join(GOLDEN, name + ".pt")
    torch.save(fixture, path)
    print("%-32s |out| max %.4f  oracle-vs-reference max|d| %.2e  -> %s (%.0f KB)" % (
        name, fixture["outputs"].abs().max().item(), pin, os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


def normalise_u8(u8, model):
    """uint8 [7,B,3,H,W] -> the 7 float frames the dataset hands the system (data/vimeo_septuplet.py:31-40,73-76)."""
    frames = [f.float() / 255 for f in u8]
    if model == "superslomo":
        mean = torch.tensor([0.429, 0.431, 0.397]).view(1, 3, 1, 1)
        frames = [f - mean for f in frames]
    elif model == "voxelflow":
        frames = [(f * 255 - 127.5) / 127.5 for f in frames]
    return frames


def structured_u8(seed, batch, h, w):
    import bench
    fr = bench.synthetic_septuplets(batch, seed, h, w)
    return torch.stack([(f * 255).round().clamp(0, 255).to(torch.uint8) for f in fr])


def apply_init_transform(named_params, kind):
    """In-place change of a seeded init BEFORE the reference runs (no reference code is touched)."""
    if kind == "sepconv_delta":
        with torch.no_grad():
            for n, p in named_params:
                if n.endswith(".7.bias") and ("moduleVertical" in n or "moduleHorizontal" in n):
                    p[p.numel() // 2] = 0.5 ** 0.5
    else:
        raise KeyError(kind)


def digest(t):
    t = t.detach().double().reshape(-1)
    return torch.tensor([t.sum(), t.abs().sum(), (t * t).sum()], dtype=torch.float64), t[:8].clone().float()


def synthetic_frames(seed, batch, size, model="sepconv"):
    """Seeded [0,1] frames, then the model's dataset normalisation (data/vimeo_septuplet.py:31-40,73-76)."""
    g = torch.Generator().manual_seed(seed)
    h, w = (size, size) if isinstance(size, int) else size
    frames = [torch.rand(batch, 3, h, w, generator=g) for _ in range(7)]
    if model == "superslomo":
        mean = torch.tensor([0.429, 0.431, 0.397]).view(1, 3, 1, 1)
        frames = [f - mean for f in frames]
    elif model == "voxelflow":
        frames = [(f * 255 - 127.5) / 127.5 for f in frames]
    return frames


def run_case(name, over, size, batch):
    from oracle import reference_shims as rs
    from oracle import maml
    over = dict(over)
    gain = over.pop("_weight_gain", None)
    vgg_seed = over.pop("_vgg_seed", None)
    init_kind = over.pop("_init", None)
    vgg_state = None
    if vgg_seed is not None:
        from oracle.super_loss import seeded_vgg16_state
        rs.install()
        import loss as ref_loss
        import torchvision.models as tv_models
        vgg_state = seeded_vgg16_state(vgg_seed)
        real_vgg16 = tv_models.vgg16

        def seeded_vgg16(pretrained=False, **kw):
            m = real_vgg16(weights=None)
            m.load_state_dict(vgg_state, strict=False)
            return m
        ref_loss.models.vgg16 = seeded_vgg16
    system, args = rs.build_system(batch_size=batch, **over)
    if gain is not None:
        with torch.no_grad():
            for p in system.net.parameters():
                if p.dim() == 4:
                    p.mul_(gain)
    if init_kind is not None:
        apply_init_transform(system.net.named_parameters(), init_kind)
    frames_u8 = None
    if isinstance(size, tuple) and size[0] == "u8":
        frames_u8 = structured_u8(size[3], batch, size[1], size[2])
        frames = normalise_u8(frames_u8, over.get("model", "sepconv"))
    else:
        frames = synthetic_frames(0, batch, size, over.get("model", "sepconv"))
    init = {k: v.detach().clone() for k, v in system.net.named_parameters()}
    att_state = {k: v.detach().clone() for k, v in system.attenuator.state_dict().items()} if args.attenuate else None

    # the oracle on the same case (before the reference mutates its parameters)
    ora = maml.OracleSystem(args.model, init, optimizer=args.optimizer, metasgd=args.metasgd,
                            num_steps=args.number_of_training_steps_per_iter, inner_lr=args.inner_lr,
                            outer_lr=args.outer_lr,
                            learnable_lr=args.learnable_per_layer_per_step_inner_loop_learning_rate, loss=args.loss,
                            attenuate=args.attenuate, use_msl=args.use_multi_step_loss_optimization,
                            msl_epochs=args.multi_step_loss_num_epochs, attenuator_state=att_state, vgg_state=vgg_state,
                            weight_decay=args.weight_decay)
    record = {}
    o_loss, o_preds, o_psnrs, o_grads = ora.run_train_iter(frames, 0, record=record)

    # capture the reference's meta-gradients by wrapping its optimizer step
    grads_ref = {}
    orig_step = system.optimizer.step

    def step_and_capture(*a, **k):
        for n, p in system.named_parameters():
            grads_ref[n] = None if p.grad is None else p.grad.detach().clone()
        return orig_step(*a, **k)

    system.optimizer.step = step_and_capture
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses, preds, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
    post = {k: v.detach().clone() for k, v in system.net.named_parameters()}

    # ---- pin: oracle == reference
    assert abs(float(losses["loss"]) - float(o_loss)) <= 1e-6 * max(1.0, abs(float(o_loss))), (name, losses["loss"], o_loss)
    for a, b in zip(preds, o_preds):
        assert (a - b).abs().max().item() <= 1e-6, name
    for k, g in o_grads["theta"].items():
        r = grads_ref["net." + k]
        assert (g - r).abs().max().item() <= 1e-6 * max(1.0, r.abs().max().item()), (name, k)
    pin = max((post[k] - ora.params[k].detach()).abs().max().item() for k in post)

    fixture = dict(
        name=name, args={k: getattr(args, k) for k in (
            "model", "loss", "optimizer", "metasgd", "attenuate", "inner_lr", "outer_lr", "batch_size", "random_seed",
            "number_of_training_steps_per_iter", "number_of_evaluation_steps_per_iter",
            "learnable_per_layer_per_step_inner_loop_learning_rate", "use_multi_step_loss_optimization",
            "multi_step_loss_num_epochs", "second_order", "first_order_to_second_order_epoch",
            "enable_inner_loop_optimizable_bn_params", "weight_decay")},
        frames=torch.stack(frames) if frames_u8 is None else None, frames_u8=frames_u8,
        loss=float(losses["loss"]), psnr=float(metrics["psnr"].avg), ssim=float(metrics["ssim"].avg),
        preds=torch.cat([p.detach() for p in preds]),
        support_losses=record.get("support_loss"),
        init_digest={k: digest(v)[0] for k, v in init.items()},
        grad_digest={k[4:]: digest(v) for k, v in grads_ref.items() if k.startswith("net.") and v is not None},
        post_digest={k: digest(v) for k, v in post.items()},
        delta_digest={k: digest(post[k] - init[k]) for k in post},      # the outer step itself (post - init)
        lr_grads={k: (None if v is None else v.clone()) for k, v in grads_ref.items()
                  if k.startswith("inner_loop_optimizer.") and v is not None and v.numel() <= 64},
        oracle_vs_reference_post_step_maxabs=pin,
        attenuator_state=att_state,
        weight_gain=gain, init_transform=init_kind,
        vgg_seed=vgg_seed,
    )
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".pt")
    torch.save(fixture, path)
    print("%-32s loss %.8f psnr %.6f  oracle-vs-reference post-step max|d| %.2e  -> %s (%.0f KB)" % (
        name, fixture["loss"], fixture["psnr"], pin, os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--test-iter", dest="test_iter", action="store_true", help="the run_test_iter fixtures")
    ap.add_argument("--full", action="store_true", help="the BASELINE-size cases (minutes of CPU each)")
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    if a.test_iter or (a.only and a.only in TEST_ITER_CASES):
        for name, (over, size, batch) in TEST_ITER_CASES.items():
            if not a.only or a.only == name:
                run_test_iter_case(name, over, size, batch)
        return
    table = dict(CASES)
    if a.full or (a.only and a.only in FULL_CASES):
        table = FULL_CASES
    for name, (over, size, batch) in table.items():
        if a.only and a.only != name:
            continue
        run_case(name, over, size, batch)


if __name__ == "__main__":
    main()
