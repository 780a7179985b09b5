"""Host-side logic (tape, plugin, inner rules, meta system, fast path) on CPU with the oracle's
operator table injected in place of the CUDA one."""
import pytest
import torch

from helpers import digest, load_golden, make_args, oracle_from_fixture, system_from_fixture
from oracle import backbones as bb


def test_plugin_parameter_names_shapes_and_seeded_init(ref_ops):
    from meta_interpolation_b200.sepconv.model import MetaNetwork
    bb.set_torch_seed(12345)
    net = MetaNetwork(ops=ref_ops)
    ref = bb.seeded_params("sepconv", 12345)
    own = dict(net.named_parameters())
    assert list(own) == list(ref)
    for k in ref:
        assert own[k].shape == ref[k].shape and torch.equal(own[k].detach(), ref[k]), k
    # checkpoint round trip keeps the reference's keys
    sd = net.state_dict()
    assert set(sd) == set(ref)
    net2 = MetaNetwork(ops=ref_ops)
    net2.load_state_dict(sd)
    assert torch.equal(net2.arena.flat, net.arena.flat)


def test_plugin_forward_and_unrouted_grads_are_none_after_step0(ref_ops):
    from meta_interpolation_b200.sepconv.model import MetaNetwork
    bb.set_torch_seed(12345)
    net = MetaNetwork(ops=ref_ops)
    own = dict(net.named_parameters())
    g = torch.Generator().manual_seed(0)
    f0, f1, tgt = (torch.rand(1, 3, 32, 32, generator=g) for _ in range(3))
    fast = {k: v.detach().clone().requires_grad_(True) for k, v in own.items()}
    out = net.forward(f0, f1, params=fast)
    assert out.shape == (1, 3, 32, 32)
    grads = torch.autograd.grad((out - tgt).abs().mean(), list(fast.values()), allow_unused=True)
    for (k, _), gr in zip(fast.items(), grads):
        assert (gr is not None) == net.is_routed(k), k
    ref = {k: v.detach().clone().requires_grad_(True) for k, v in own.items()}
    o2 = bb.sepconv_forward(f0, f1, ref, ref)
    assert (out - o2).abs().max().item() < 1e-5


@pytest.mark.parametrize("name,fast", [("sepconv_lslr_sgd_k2", True), ("sepconv_lslr_sgd_k2", False),
                                       ("sepconv_lslr_sgd_k1_b2_mse", True),
                                       ("sepconv_lslr_learnable_msl_k2", True),
                                       ("sepconv_lslr_learnable_msl_k2", False),
                                       ("sepconv_lslr_adam_k2", False), ("sepconv_metasgd_adamax_k2", False),
                                       ("sepconv_lslr_adam_k2", True), ("sepconv_metasgd_adamax_k2", True),
                                       ("sepconv_l2f_sgd_k1", False), ("sepconv_l2f_sgd_k1", True)])
def test_system_against_reference_golden(ref_ops, name, fast):
    fx = load_golden(name)
    system = system_from_fixture(fx, ref_ops, fast_path=fast)
    assert system.fast_path_supported() == fast
    frames = list(fx["frames"])
    losses, preds, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
    assert abs(float(losses["loss"]) - fx["loss"]) <= 2e-6
    assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 2e-6
    assert abs(metrics["psnr"].avg - fx["psnr"]) < 0.01          # north_star tolerance: |dPSNR| < 0.01 dB
    if fx["args"]["optimizer"] == "SGD":                          # Adam/Adamax steps are sign-like: digest only for SGD
        own = dict(system.net.named_parameters())
        for k, (d, head) in fx["post_digest"].items():
            assert torch.allclose(digest(own[k])[0], d, rtol=1e-5, atol=1e-8), k


FLOW_SIZES = {"voxelflow": (72, 88), "superslomo": (72, 80), "rrin": (72, 136), "cain": (120, 136)}
FLOW_COUNTS = {"voxelflow": 23, "superslomo": 92, "rrin": 162, "cain": 494}


@pytest.mark.parametrize("model", ["voxelflow", "superslomo", "rrin", "cain"])
def test_flow_plugins_names_init_forward_and_routing(ref_ops, model):
    """Parameter schema + seeded init == reference; plugin forward/backward == oracle forward under autograd;
    tensors the reference never routes come back with a None gradient."""
    from meta_interpolation_b200.meta_learning_system import _build_backbone
    bb.set_torch_seed(12345)
    net = _build_backbone(make_args(model=model), ref_ops)
    ref = bb.seeded_params(model, 12345)
    own = dict(net.named_parameters())
    assert list(own) == list(ref) and len(own) == FLOW_COUNTS[model]
    for k in ref:
        assert own[k].shape == ref[k].shape and torch.equal(own[k].detach(), ref[k]), k
    g = torch.Generator().manual_seed(0)
    f0, f1, tgt = (torch.rand(1, 3, *FLOW_SIZES[model], generator=g) for _ in range(3))
    # cain's 125 stacked xavier convs explode at the default init (|out| ~ 1e2) and make the gradient
    # comparison ill-conditioned: its conv weights are scaled down for this check
    gain = 0.4 if model == "cain" else 1.0
    fast = {k: (v.detach().clone() * (gain if v.dim() == 4 else 1.0)).requires_grad_(True) for k, v in own.items()}
    out = net.forward(f0, f1, params=fast)
    if model == "superslomo":
        out, extras = out
        assert set(extras) == {"bidirectional_flow", "warped_intermediate_frames", "warped_input_frames"}
        assert all(t.shape[2:] == f0.shape[2:] for pair in extras.values() for t in pair)
    assert out.shape == f0.shape
    grads = torch.autograd.grad(((out - tgt) ** 2).mean(), list(fast.values()), allow_unused=True)
    fr = {k: v.detach().clone().requires_grad_(True) for k, v in fast.items()}
    o2 = bb.BACKBONES[model]["forward"](f0, f1, fr, {k: v.detach() for k, v in own.items()})
    g2 = torch.autograd.grad(((o2 - tgt) ** 2).mean(), list(fr.values()), allow_unused=True)
    scale = max(1.0, o2.abs().max().item())
    assert (out - o2).abs().max().item() <= 1e-5 * scale
    routed = bb.BACKBONES[model]["is_routed"]
    gmax = max(b.abs().max().item() for b in g2 if b is not None)
    for (k, _), a, b in zip(fast.items(), grads, g2):
        assert (a is not None) == routed(k) == net.is_routed(k), k
        assert (a is None) == (b is None), k
        if a is not None:   # cain at its default init is badly conditioned: compare on the global gradient scale
            assert (a - b).abs().max().item() <= 2e-3 * max(b.abs().max().item(), 1e-3 * gmax), k


FLOW_GOLDEN = ["voxelflow_lslr_sgd_k1_mse", "voxelflow_lslr_sgd_k2_ragged", "superslomo_metasgd_sgd_k2",
               "superslomo_lslr_sgd_k1_ragged", "rrin_msl_learnable_k2", "rrin_lslr_sgd_k1_ragged",
               "cain_l2f_sgd_k1", "cain_lslr_sgd_k2_ragged", "cain_lslr_sgd_k2_gain04", "cain_l2f_sgd_k1_gain04",
               "voxelflow_metasgd_adam_k1"]


@pytest.mark.parametrize("name", FLOW_GOLDEN)
@pytest.mark.parametrize("fast", [True, False])
def test_flow_systems_against_reference_golden(ref_ops, name, fast):
    """BASELINE configs[0] (voxelflow 128x128 K=1) and configs[2..4] in miniature against the reference's outputs."""
    fx = load_golden(name)
    system = system_from_fixture(fx, ref_ops, fast_path=fast)
    assert system.fast_path_supported() == fast     # every configuration here, L2F included, has a graph path
    frames = list(fx["frames"])
    losses, preds, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
    scale = max(1.0, fx["preds"].abs().max().item())           # cain's default init explodes (|pred| ~ 1e2)
    assert abs(float(losses["loss"]) - fx["loss"]) <= 1e-5 * max(1.0, abs(fx["loss"]))
    assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 2e-4 * scale
    assert abs(metrics["psnr"].avg - fx["psnr"]) < 0.01          # north_star tolerance
    if abs(fx["loss"]) < 5:
        assert abs(float(metrics["ssim"].avg) - fx["ssim"]) < 1e-4   # pytorch_msssim restatement (utils.ssim)
    own = dict(system.net.named_parameters())
    tol = 0.2 if fx["args"]["model"] == "cain" else 5e-3
    for k, (d, head) in fx["post_digest"].items():
        mine = digest(own[k])[0]
        assert torch.allclose(mine[1:], d[1:], rtol=tol, atol=1e-8), k
    if fx.get("delta_digest") is not None:
        # the outer step itself (post - init): for voxelflow + Adam it is the policy-built Adam with weight decay
        # (reference :133-136), a sign-like step of outer_lr per element
        init = oracle_from_fixture(fx).params
        for k, (d, head) in fx["delta_digest"].items():
            mine = digest(own[k].detach() - init[k].detach())[0]
            assert torch.allclose(mine[1:], d[1:], rtol=2e-2, atol=1e-12), k


def test_voxelflow_adam_optimizer_follows_the_reference_policies(ref_ops):
    """reference :133-136 + voxel_flow.py:307-350: three param groups (conv weights / conv bias / BN scale+shift)
    over the backbone's tensors only, weight_decay from args, torch's default betas; Meta-SGD alphas are NOT stepped."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    s = SceneAdaptiveInterpolation(make_args(model="voxelflow", optimizer="Adam", metasgd=True, loss="1*MSE"),
                                   ops=ref_ops)
    sd = s.optimizer.state_dict()
    assert [g["name"] for g in sd["param_groups"]] == ["model weight", "model bias", "model bn scale/shift"]
    assert [len(g["params"]) for g in sd["param_groups"]] == [8, 1, 14]
    assert [(g["lr_mult"], g["decay_mult"]) for g in sd["param_groups"]] == [(1, 1), (2, 0), (1, 1)]
    assert all(g["betas"] == (0.9, 0.999) and g["weight_decay"] == 1e-4 for g in sd["param_groups"])
    assert s.fast_path_supported()
    alpha0 = s.alpha.flat.clone()
    g = torch.Generator().manual_seed(2)
    frames = [torch.rand(1, 3, 64, 64, generator=g) * 2 - 1 for _ in range(7)]
    s.run_train_iter(frames, epoch=0)
    assert torch.equal(s.alpha.flat, alpha0)


def test_state_dict_keys_match_reference_schema(ref_ops):
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    s = SceneAdaptiveInterpolation(make_args(attenuate=True, number_of_training_steps_per_iter=2), ops=ref_ops)
    keys = list(s.state_dict().keys())
    assert "net.moduleConv1.0.weight" in keys and "gamma_mult" in keys
    assert "inner_loop_optimizer.names_learning_rates_dict.moduleConv1-0-weight" in keys
    assert s.state_dict()["inner_loop_optimizer.names_learning_rates_dict.moduleConv1-0-weight"].shape == (3,)
    assert {"attenuator.0.weight", "attenuator.0.bias", "attenuator.2.weight", "attenuator.2.bias"} <= set(keys)
    assert len(keys) == 94 + 94 + 4 + 1     # 193 keys for sepconv+L2F


def test_metasgd_sgd_reproduces_reference_failure(ref_ops):
    # Meta-SGD + SGD with K>=2 on a backbone with un-routed tensors raises in the reference
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    s = SceneAdaptiveInterpolation(make_args(metasgd=True, number_of_training_steps_per_iter=2), ops=ref_ops)
    g = torch.Generator().manual_seed(0)
    frames = [torch.rand(1, 3, 32, 32, generator=g) for _ in range(7)]
    with pytest.raises(TypeError):
        s.run_train_iter(frames, epoch=0)


def test_validation_and_test_iters(ref_ops):
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    from oracle import maml
    s = SceneAdaptiveInterpolation(make_args(number_of_evaluation_steps_per_iter=1), ops=ref_ops)
    g = torch.Generator().manual_seed(3)
    frames = [torch.rand(1, 3, 32, 32, generator=g) for _ in range(7)]
    losses, preds, metrics = s.run_validation_iter(frames)
    ora = maml.OracleSystem("sepconv", bb.seeded_params("sepconv", 12345), num_steps=1)
    loss, opreds, psnrs = ora.run_validation_iter(frames, num_steps=1)
    assert abs(float(losses["loss"]) - float(loss)) < 2e-6
    assert (preds[0] - opreds[0]).abs().max().item() < 2e-6
    out = s.run_test_iter(frames[:4])
    assert out[0].shape == (3, 32, 32)


def test_extract_top_level_dict():
    from meta_interpolation_b200.model_utils import extract_top_level_dict
    d = {"a.0.weight": 1, "a.0.bias": 2, "b.weight": 3, "c": 4, "layer_dict.e.f": 5}
    out = extract_top_level_dict(d)
    assert out == {"a": {"0.weight": 1, "0.bias": 2}, "b": {"weight": 3}, "c": 4, "e": {"f": 5}}


@pytest.mark.parametrize("name", ["sepconv_l2f_sgd_k1"])
def test_l2f_graph_path_matches_compat_path_on_attenuator_gradients(ref_ops, name):
    """The goldens digest only the backbone parameters; the L2F-specific outer gradients (gamma_mult, attenuator
    MLP; reference meta_learning_system.py:107-117, 258-272) are checked path against path: the graph path derives
    them from dL/dgamma_i = <G_i, theta_i>, the compat path from autograd through the whole inner loop."""
    fx = load_golden(name)
    delta = {}
    for fast in (False, True):
        system = system_from_fixture(fx, ref_ops, fast_path=fast)
        with torch.no_grad():
            system.gamma_mult.fill_(0.3)      # away from the init (0), where the attenuator gradients vanish
        system.optimizer.param_groups[0]["lr"] = 1.0    # SGD outer step: the parameter change IS minus the gradient
        pick = lambda: {k: v.detach().clone() for k, v in system.state_dict().items()
                        if k.startswith("attenuator") or k == "gamma_mult"}
        before = pick()
        system.run_train_iter(list(fx["frames"]), epoch=0)
        after = pick()
        delta[fast] = {k: after[k] - before[k] for k in before}
    assert delta[True]["gamma_mult"].abs().item() > 0
    for k in delta[False]:
        a, b = delta[True][k], delta[False][k]
        assert b.abs().max().item() > 0, k                         # every L2F tensor receives a gradient
        assert (a - b).abs().max().item() <= 1e-3 * b.abs().max().item(), (k, (a - b).abs().max().item())


@pytest.mark.parametrize("name", ["sepconv_lslr_adam_k2", "sepconv_metasgd_adamax_k2"])
def test_moment_rules_graph_path_matches_compat_path_on_outer_gradients(ref_ops, name):
    """Adam / Adamax inner rules (reference inner_loop_optimizers.py:150-244, :335-426): the goldens pin loss and
    predictions; the outer gradients (theta, and Meta-SGD's alpha = -((theta - w_K)/alpha) (.) G on the graph path)
    are checked path against path on the flat gradient buffers handed to the outer step."""
    fx = load_golden(name)
    grads = {}
    for fast in (False, True):
        system = system_from_fixture(fx, ref_ops, fast_path=fast)
        assert system.fast_path_supported() == fast
        opt, seen = system.optimizer, {}
        orig = opt.step

        def step(opt=opt, seen=seen, orig=orig):
            opt.gather_grads()
            for g in opt.flat_groups:
                seen[g.name] = g.grad.detach().clone()
            orig()
        opt.step = step
        system.run_train_iter(list(fx["frames"]), epoch=0)
        grads[fast] = seen
    assert set(grads[True]) == set(grads[False]) and len(grads[True]) == (2 if fx["args"]["metasgd"] else 1)
    for k, b in grads[False].items():
        a = grads[True][k]
        assert b.abs().max().item() > 0, k
        tol = 1e-4 * b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() <= tol, (k, (a - b).abs().max().item(), b.abs().max().item())


@pytest.mark.parametrize("kw", [
    dict(optimizer="Adamax", metasgd=True, use_multi_step_loss_optimization=True, multi_step_loss_num_epochs=5,
         inner_lr=1e-4, number_of_training_steps_per_iter=2),
    dict(optimizer="Adamax", inner_lr=1e-7, number_of_training_steps_per_iter=3),
    dict(optimizer="Adam", inner_lr=1e-5, number_of_training_steps_per_iter=2,
         learnable_per_layer_per_step_inner_loop_learning_rate=True, eval_first=True),
    dict(optimizer="SGD", inner_lr=1e-5, number_of_training_steps_per_iter=2,
         learnable_per_layer_per_step_inner_loop_learning_rate=True, eval_first=True)],
    ids=["metasgd_adamax_msl", "lslr_adamax_k3", "learnable_lr_adam", "learnable_lr_sgd_eval_first"])
def test_moment_rule_combinations_graph_path_matches_compat_path(ref_ops, kw):
    """Meta-SGD-Adamax under the multi-step loss (alpha gradient per step from (theta - w_k)/alpha) and the
    LSLR-Adamax quirk over three steps (exp_avg persists, exp_inf does not; inner_loop_optimizers.py:229-236; tiny lr:
    that rule is chaotic at practical rates in the reference too); learnable per-step rates under Adam (lr gradient
    from the stored weight deltas); train and validation iterations, validation FIRST in the learnable-lr cases (a
    ``--mode val`` run evaluates before any training iteration has sized the per-step buffers)."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    g = torch.Generator().manual_seed(4)
    frames = [torch.rand(1, 3, 32, 40, generator=g) for _ in range(7)]
    res = {}
    kw = dict(kw)
    eval_first = kw.pop("eval_first", False)
    for fast in (True, False):
        s = SceneAdaptiveInterpolation(make_args(number_of_evaluation_steps_per_iter=2, fast_path=fast, **kw),
                                       ops=ref_ops)
        assert s.fast_path_supported() == fast
        if eval_first:
            s.run_validation_iter(frames)
        opt, seen = s.optimizer, {}
        orig = opt.step

        def step(opt=opt, seen=seen, orig=orig):
            opt.gather_grads()
            for gr in opt.flat_groups:
                seen[gr.name] = gr.grad.detach().clone()
            orig()
        opt.step = step
        lt, pt, _ = s.run_train_iter(frames, epoch=0)
        lv, pv, _ = s.run_validation_iter(frames)
        res[fast] = (float(lt["loss"].detach()), torch.cat(pt), seen, float(lv["loss"].detach()), torch.cat(pv))
    a, b = res[True], res[False]
    assert abs(a[0] - b[0]) <= 2e-6 and abs(a[3] - b[3]) <= 2e-6
    assert (a[1] - b[1]).abs().max().item() <= 5e-6
    assert (a[4] - b[4]).abs().max().item() <= 5e-5      # after a sign-like Adamax OUTER step on both paths
    for k, gb in b[2].items():
        # (the LSLR-Adamax direction m/(|g|+eps) magnifies fp32 summation-order noise wherever |g| ~ eps)
        assert (a[2][k] - gb).abs().max().item() <= 1e-3 * gb.abs().max().item() + 1e-12, k


@pytest.mark.parametrize("model,hw", [("sepconv", (32, 40)), ("superslomo", (64, 64))])
def test_test_time_adaptation_graph_path_matches_compat_path(ref_ops, model, hw):
    """run_test_iter (reference :630-697: 4-frame clips, support (0,1,2),(1,2,3), query (1,2)) through the captured
    fast path equals the compat transcription of the reference's control flow."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    g = torch.Generator().manual_seed(11)
    frames = [torch.rand(2, 3, *hw, generator=g) - (0.4 if model == "superslomo" else 0.0) for _ in range(4)]
    outs = {}
    for fast in (True, False):
        s = SceneAdaptiveInterpolation(make_args(model=model, number_of_evaluation_steps_per_iter=2, fast_path=fast),
                                       ops=ref_ops)
        assert s.fast_path_supported() == fast
        outs[fast] = s.run_test_iter(frames)
    for a, b in zip(outs[True], outs[False]):
        assert a.shape == (3,) + hw and (a - b).abs().max().item() <= 2e-6


def test_tiled_evaluation_follows_experiment_builder_splitting(ref_ops, monkeypatch):
    """Frames above 5e5 pixels are halved along the longer side (experiment_builder.py:101-128, 153-172)."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    s = SceneAdaptiveInterpolation(make_args(), ops=ref_ops)
    calls = []

    def fake_val(frames):
        calls.append(tuple(frames[0].shape[-2:]))
        return {"loss": torch.tensor(float(len(calls)))}, [f.clone() for f in frames[3]], None

    monkeypatch.setattr(s, "run_validation_iter", fake_val)
    frames = [torch.arange(2 * 3 * 8 * 6, dtype=torch.float32).view(2, 3, 8, 6) + i for i in range(7)]
    monkeypatch.setattr(s, "_needs_tiling", lambda fr: fr[0].shape[-2] * fr[0].shape[-1] > 20)
    losses, outs = s.run_validation_iter_tiled(frames)
    # 8x6=48 > 20 -> halve rows (H > W): 4x6=24 > 20 -> halve columns (W > H): 4x3 = 12
    assert calls == [(4, 3)] * 4
    assert torch.equal(torch.stack(outs), frames[3])             # stitched back in place
    assert float(losses["loss"]) == ((1 + 2) / 2 + (3 + 4) / 2) / 2


@pytest.mark.parametrize("opt", ["Adam", "Adamax", "SGD"])
def test_outer_optimizer_state_dict_round_trip(ref_ops, opt):
    """The fused outer optimizer writes / reads the stock per-parameter ``torch.optim`` state layout: a fresh system
    restored from (model state_dict, optimizer state_dict) continues bit-identically to the one that kept running."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    kw = dict(optimizer=opt, outer_lr=1e-3, number_of_training_steps_per_iter=1, metasgd=(opt == "Adamax"))
    g = torch.Generator().manual_seed(2)
    batches = [[torch.rand(1, 3, 32, 32, generator=g) for _ in range(7)] for _ in range(2)]
    a = SceneAdaptiveInterpolation(make_args(**kw), ops=ref_ops)
    a.run_train_iter(batches[0], epoch=0)
    sd_model = {k: v.detach().clone() for k, v in a.state_dict().items()}
    sd_opt = a.optimizer.state_dict()
    n_params = len(list(a.trainable_parameters()))
    assert sd_opt["param_groups"][0]["params"] == list(range(n_params))
    if opt == "SGD":
        assert sd_opt["state"] == {}
    else:
        second = "exp_avg_sq" if opt == "Adam" else "exp_inf"
        assert len(sd_opt["state"]) == n_params
        for i, p in enumerate(a.trainable_parameters()):
            assert sd_opt["state"][i]["exp_avg"].shape == p.shape and sd_opt["state"][i][second].shape == p.shape
            assert sd_opt["state"][i]["exp_avg"].is_contiguous()
    b = SceneAdaptiveInterpolation(make_args(**kw), ops=ref_ops)
    b.load_state_dict(sd_model)
    b.optimizer.load_state_dict(sd_opt)
    a.run_train_iter(batches[1], epoch=0)
    b.run_train_iter(batches[1], epoch=0)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k


def test_super_loss_system_against_reference_golden(ref_ops):
    """`--loss 1*Super` (loss.py:246-274, scripts/run_superslomo.sh): one meta-iteration of the graph path against the
    golden generated by the unmodified reference with a seeded random VGG16 conv4_3 on both sides (the ImageNet
    weights are not available offline); loss, predictions, PSNR and the post-step parameters."""
    from oracle.super_loss import seeded_vgg16_state
    fx = load_golden("superslomo_super_sgd_k1")
    system = system_from_fixture(fx, ref_ops, fast_path=True, vgg16_weights=seeded_vgg16_state(fx["vgg_seed"]))
    assert system.fast_path_supported()
    grads, orig = {}, system.optimizer.step

    def step():
        for k in fx["grad_digest"]:
            grads[k] = system.net_grad.reference_view(k).detach().clone()
        orig()
    system.optimizer.step = step
    losses, preds, metrics = system.run_train_iter(list(fx["frames"]), epoch=0, do_evaluation=True)
    assert abs(float(losses["loss"]) - fx["loss"]) <= 2e-6 * fx["loss"]
    assert float(losses["Super"]) == float(losses["total"])
    assert len(grads) == 92                                   # every SuperSloMo tensor receives a meta-gradient
    for k, (d, head) in fx["grad_digest"].items():            # the reference's own outer gradients
        mine = digest(grads[k])[0]
        # (sum, sum|.|, sum of squares): fp32 summation-order noise, measured against the absolute sum
        assert abs(float(mine[0] - d[0])) <= 3e-4 * float(d[1]) and abs(float(mine[1] - d[1])) <= 3e-4 * float(d[1]), \
            (k, mine, d)
        assert abs(float(mine[2] - d[2])) <= 1e-3 * float(d[2]), (k, mine, d)
    assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 2e-6
    assert abs(metrics["psnr"].avg - fx["psnr"]) < 0.01
    own = dict(system.net.named_parameters())
    for k, (d, head) in fx["post_digest"].items():
        assert torch.allclose(digest(own[k])[0], d, rtol=1e-5, atol=1e-8), k
    # validation (forward-only query) reports the same loss the training query of an identical system would
    lv, _, _ = system.run_validation_iter(list(fx["frames"]))
    assert torch.isfinite(lv["loss"]).item()
    # the compat path (autograd over one fused forward+loss step) reproduces the same golden
    compat = system_from_fixture(fx, ref_ops, fast_path=False, vgg16_weights=seeded_vgg16_state(fx["vgg_seed"]))
    assert not compat.fast_path_supported()
    losses, preds, metrics = compat.run_train_iter(list(fx["frames"]), epoch=0, do_evaluation=True)
    assert abs(float(losses["loss"]) - fx["loss"]) <= 2e-6 * fx["loss"]
    assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 2e-6
    own = dict(compat.net.named_parameters())
    for k, (d, head) in fx["post_digest"].items():
        assert torch.allclose(digest(own[k])[0], d, rtol=1e-5, atol=1e-8), k


@pytest.mark.parametrize("kw,hw", [
    (dict(model="superslomo", metasgd=True, number_of_training_steps_per_iter=2), (64, 64)),
    (dict(model="sepconv", learnable_per_layer_per_step_inner_loop_learning_rate=True,
          use_multi_step_loss_optimization=True, multi_step_loss_num_epochs=5, number_of_training_steps_per_iter=2),
     (32, 40))], ids=["l2f_metasgd", "l2f_learnable_lr_msl"])
def test_l2f_combinations_graph_path_matches_compat_path(ref_ops, kw, hw):
    """L2F (--attenuate) combined with Meta-SGD, and with learnable per-step rates under the multi-step loss: the graph
    path's outer gradients (theta through gamma, alpha / lr from the stored query gradient, attenuator and gamma_mult
    from dL/dgamma_i = <G_i, theta_i> per query pass) against autograd through the compat path."""
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    g = torch.Generator().manual_seed(6)
    shift = 0.4 if kw["model"] == "superslomo" else 0.0
    frames = [torch.rand(1, 3, *hw, generator=g) - shift for _ in range(7)]
    res = {}
    for fast in (True, False):
        s = SceneAdaptiveInterpolation(make_args(attenuate=True, inner_lr=1e-4, fast_path=fast, **kw), ops=ref_ops)
        assert s.fast_path_supported() == fast
        with torch.no_grad():
            s.gamma_mult.fill_(0.3)
        opt, seen = s.optimizer, {}
        orig = opt.step

        def step(opt=opt, seen=seen, orig=orig):
            opt.gather_grads()
            for gr in opt.flat_groups:
                seen[gr.name] = gr.grad.detach().clone()
            orig()
        opt.step = step
        lt, pt, _ = s.run_train_iter(frames, epoch=0)
        res[fast] = (float(lt["loss"].detach()), torch.cat(pt), seen)
    a, b = res[True], res[False]
    assert abs(a[0] - b[0]) <= 2e-6 and (a[1] - b[1]).abs().max().item() <= 5e-6
    assert len(b[2]) == 3                                   # theta, alpha | lr, and the L2F group
    for k, gb in b[2].items():
        assert gb.abs().max().item() > 0, k
        assert (a[2][k] - gb).abs().max().item() <= 1e-4 * gb.abs().max().item() + 1e-12, k


TEST_ITER_GOLDEN = ["test_iter_sepconv_k2", "test_iter_superslomo_k2", "test_iter_voxelflow_k1", "test_iter_rrin_l2f_k1"]


def system_for_test_iter(fx, ops, fast):
    s = system_from_fixture(fx, ops, fast_path=fast)
    if fx.get("gamma_mult") is not None:
        with torch.no_grad():
            s.gamma_mult.fill_(fx["gamma_mult"])
    return s


@pytest.mark.parametrize("name", TEST_ITER_GOLDEN)
@pytest.mark.parametrize("fast", [True, False])
def test_run_test_iter_against_reference_golden(ref_ops, name, fast):
    """run_test_iter (reference meta_learning_system.py:630-697) against what the UNMODIFIED reference returned for
    the same 4-frame clips (oracle/make_golden.py --test-iter): only superslomo is de-normalised, L2F attenuates from
    the test-mode support triplets."""
    fx = load_golden(name)
    s = system_for_test_iter(fx, ref_ops, fast)
    assert s.fast_path_supported() == fast
    outs = s.run_test_iter(list(fx["frames"]))
    assert len(outs) == fx["outputs"].shape[0]
    for a, b in zip(outs, fx["outputs"]):
        assert a.shape == b.shape and (a - b).abs().max().item() <= 1e-5


@pytest.mark.parametrize("name", TEST_ITER_GOLDEN)
def test_oracle_run_test_iter_reproduces_reference_golden(name):
    fx = load_golden(name)
    ora = oracle_from_fixture(fx)
    ora.num_steps = fx["args"]["number_of_evaluation_steps_per_iter"]
    if fx.get("gamma_mult") is not None:
        with torch.no_grad():
            ora.gamma_mult.fill_(fx["gamma_mult"])
    outs = ora.run_test_iter(list(fx["frames"]))
    for a, b in zip(outs, fx["outputs"]):
        assert (a - b).abs().max().item() <= 1e-6


@pytest.mark.parametrize("name,fast", [("sepconv_lslr_sgd_k2", True), ("sepconv_lslr_sgd_k2", False),
                                       ("sepconv_metasgd_adamax_k2", True), ("sepconv_l2f_sgd_k1", True),
                                       ("cain_lslr_sgd_k2_gain04", True)])
def test_tf32_operand_convention_host_logic(name, fast):
    """The TF32 operand convention (include/mi_b200.h): fprop reads TF32-rounded weight copies (meta shadow arena,
    per-lane fast shadow written by the fused update / after the flat moment update / after the L2F attenuation),
    dgrad rounded rotated copies, activations and gradients are rounded before a conv reads them, and the exact fp32
    master weights are what the updates see.  Run with an operator table that rounds like the GPU but convolves
    exactly: results must stay within the rounding noise of the golden (and the masters must NOT be on the grid)."""
    from oracle.ops_ref import RefOps
    from meta_interpolation_b200 import backbone
    ops = RefOps(tf32_rn=True)
    saved = backbone._default_ops
    backbone.set_default_ops(ops)
    try:
        fx = load_golden(name)
        system = system_from_fixture(fx, ops, fast_path=fast)
        assert system.fast_path_supported() == fast
        losses, preds, metrics = system.run_train_iter(list(fx["frames"]), epoch=0, do_evaluation=True)
        assert abs(float(losses["loss"]) - fx["loss"]) <= 5e-4 * max(1.0, abs(fx["loss"]))
        assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 2e-3 * max(1.0, fx["preds"].abs().max().item())
        assert abs(metrics["psnr"].avg - fx["psnr"]) < 0.01
        flat = system.net.arena.flat
        grid = ops.round_tf32(flat.clone())
        assert (grid != flat).any()                       # the master copy keeps its low mantissa bits
        if fast:
            fp = system.fast_path()
            assert torch.equal(ops.round_tf32(fp.meta_r.flat.clone()), fp.meta_r.flat)      # the shadow is on the grid
            lane = fp.lanes[0]
            for n in system.net.conv_names:
                if system.net.is_routed(n + ".weight"):
                    w, wr = lane.fast.kernel_view(n + ".weight"), lane.fast_r.kernel_view(n + ".weight")
                    assert torch.equal(ops.round_tf32(w.clone().contiguous()), wr.contiguous()), n
    finally:
        backbone.set_default_ops(saved)


def test_program_cache_is_lru_bounded(ref_ops):
    """ADVICE (round 1): captured programs were cached forever per lane -- one set per frame size, accumulation scale
    and SM share.  The cache now evicts the least recently used entry."""
    fx = load_golden("sepconv_lslr_sgd_k2")
    system = system_from_fixture(fx, ref_ops, fast_path=True)
    fp = system.fast_path()
    lane = fp.lanes[0]
    fp.MAX_PROGRAMS_PER_LANE = 2
    body = lambda prog: None
    a = fp._program(lane, ('query', 'a'), body, 1, 8, 8)
    b = fp._program(lane, ('query', 'b'), body, 1, 8, 8)
    assert fp._program(lane, ('query', 'a'), body, 1, 8, 8) is a          # a hit refreshes the entry ...
    c = fp._program(lane, ('query', 'c'), body, 1, 8, 8)
    keys = [k[:2] for k in lane.programs]
    assert keys == [('query', 'a'), ('query', 'c')] and len(lane.programs) == 2   # ... so b, not a, was evicted
    assert fp._program(lane, ('query', 'b'), body, 1, 8, 8) is not b
    assert c is lane.programs[('query', 'c', fp.sm_budget)]
