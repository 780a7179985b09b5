"""The oracle against its pins: literal kernel-loop transcription, closed forms, golden vectors."""
import os

import pytest
import torch

from helpers import golden_names, load_golden, oracle_from_fixture, digest
from oracle import backbones as bb, inner_rules, maml, sepconv_op


def test_sepconv_vectorised_matches_kernel_loops():
    torch.manual_seed(0)
    f = 5
    inp = torch.rand(2, 3, 4 + f - 1, 3 + f - 1, dtype=torch.float64)
    v = torch.rand(2, f, 4, 3, dtype=torch.float64)
    h = torch.rand(2, f, 4, 3, dtype=torch.float64)
    go = torch.rand(2, 3, 4, 3, dtype=torch.float64)
    assert (sepconv_op.sepconv_forward(inp, v, h) - sepconv_op.sepconv_forward_loops(inp, v, h)).abs().max() < 1e-13
    gv, gh = sepconv_op.sepconv_backward(inp, v, h, go)
    gv2, gh2 = sepconv_op.sepconv_backward_loops(inp, v, h, go)
    assert (gv - gv2).abs().max() < 1e-13 and (gh - gh2).abs().max() < 1e-13


def test_sepconv_backward_is_the_autograd_gradient():
    torch.manual_seed(1)
    f = 3
    inp = torch.rand(1, 2, 6, 7, dtype=torch.float64)
    v = torch.rand(1, f, 4, 5, dtype=torch.float64, requires_grad=True)
    h = torch.rand(1, f, 4, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(sepconv_op.FunctionSepconvCPU.apply, (inp, v, h))


def test_delta_kernel_is_a_crop():
    # KAT: one-hot filters at the centre tap reproduce the centre crop
    f = 5
    inp = torch.rand(1, 3, 8 + f - 1, 9 + f - 1)
    v = torch.zeros(1, f, 8, 9)
    h = torch.zeros(1, f, 8, 9)
    v[:, f // 2] = 1
    h[:, f // 2] = 1
    out = sepconv_op.sepconv_forward(inp, v, h)
    assert torch.equal(out, inp[:, :, f // 2:f // 2 + 8, f // 2:f // 2 + 9])


def test_msl_weights_closed_form():
    # epoch 0 -> uniform 1/K (SURVEY Appx E6)
    assert torch.allclose(maml.msl_weights(5, 0, 1), torch.full((5,), 0.2))
    w = maml.msl_weights(5, 1, 1)
    assert torch.allclose(w[:4], torch.full((4,), 0.03 / 5)) and abs(float(w[4]) - (1 - 4 * 0.03 / 5)) < 1e-7
    assert torch.equal(maml.msl_weights(0, 3, 1), torch.ones(1))


def test_zero_gradient_update_is_identity_and_none_is_dropped():
    w = {"a.weight": torch.rand(3), "b.weight": torch.rand(2)}
    lrs = {"a-weight": torch.full((3,), 0.1), "b-weight": torch.full((3,), 0.1)}
    out = inner_rules.update_params("SGD", False, w, {"a.weight": torch.zeros(3), "b.weight": None}, lrs, {}, 0)
    assert list(out) == ["a.weight"] and torch.equal(out["a.weight"], w["a.weight"])
    with pytest.raises(TypeError):
        inner_rules.update_params("SGD", True, w, {"a.weight": None}, {"a-weight": torch.ones(3)}, {}, 0)


def test_adamax_quirks():
    # LSLR-Adamax: exp_avg persists, exp_inf does not (inner_loop_optimizers.py:229-236)
    w = {"a": torch.ones(4)}
    g = torch.tensor([0.5, -0.25, 2.0, 1e-3])
    lrs = {"a": torch.full((3,), 0.1)}
    st = {}
    o1 = inner_rules.update_params("Adamax", False, w, {"a": g}, lrs, st, 0)
    m1 = 0.1 * g
    assert torch.allclose(o1["a"], w["a"] - (0.1 / (1 - 0.9)) * m1 / (g.abs() + 1e-8))
    o2 = inner_rules.update_params("Adamax", False, o1, {"a": g}, lrs, st, 1)
    m2 = 0.9 * m1 + 0.1 * g
    assert torch.allclose(o2["a"], o1["a"] - (0.1 / (1 - 0.81)) * m2 / (g.abs() + 1e-8))
    # Meta-SGD-Adamax: stateless (:409,:418)
    st = {}
    lrs = {"a": torch.full((4,), 0.1)}
    p1 = inner_rules.update_params("Adamax", True, w, {"a": g}, lrs, st, 0)
    p2 = inner_rules.update_params("Adamax", True, p1, {"a": g}, lrs, st, 1)
    assert torch.allclose(p2["a"], p1["a"] - (0.1 / (1 - 0.81)) * (0.1 * g) / (g.abs() + 1e-8))


def test_psnr_formula():
    a = torch.full((3, 4, 4), 0.5)
    b = torch.full((3, 4, 4), 0.5 + 2 / 255)
    import math
    assert abs(maml.psnr(a, b) - (-10 * math.log10((2 / 255) ** 2 + 1e-8))) < 1e-6
    assert abs(maml.psnr(a, a) - 80.0) < 1e-6


def test_seeded_init_digest_matches_reference_fixture():
    fx = load_golden("sepconv_lslr_sgd_k2")
    params = bb.seeded_params("sepconv", fx["args"]["random_seed"])
    assert list(params) == list(fx["init_digest"])
    for k, v in params.items():
        assert torch.equal(digest(v)[0], fx["init_digest"][k]), k


@pytest.mark.parametrize("name", ["sepconv_lslr_sgd_k2", "sepconv_lslr_learnable_msl_k2", "sepconv_metasgd_adamax_k2",
                                  "sepconv_l2f_sgd_k1", "voxelflow_lslr_sgd_k1_mse", "superslomo_metasgd_sgd_k2",
                                  "rrin_msl_learnable_k2", "rrin_lslr_sgd_k1_ragged", "cain_l2f_sgd_k1"])
def test_oracle_reproduces_reference_golden(name):
    """Outputs of the UNMODIFIED reference (tests/golden, made by oracle/make_golden.py) vs the oracle."""
    fx = load_golden(name)
    ora = oracle_from_fixture(fx)
    frames = list(fx["frames"])
    loss, preds, psnrs, grads = ora.run_train_iter(frames, 0)
    assert abs(float(loss) - fx["loss"]) <= 1e-6
    assert (torch.cat(preds) - fx["preds"]).abs().max().item() <= 1e-6
    assert abs(sum(psnrs) / len(psnrs) - fx["psnr"]) <= 1e-4
    for k, (d, head) in fx["grad_digest"].items():
        mine = digest(grads["theta"][k])
        assert torch.allclose(mine[0], d, rtol=1e-5, atol=1e-9), k
    for k, (d, head) in fx["post_digest"].items():
        assert torch.allclose(digest(ora.params[k])[0], d, rtol=1e-6, atol=1e-9), k


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("model,size", [("voxelflow", (40, 72)), ("superslomo", (72, 64)), ("rrin", (64, 136)),
                                        ("cain", (128, 120))])
def test_flow_oracles_equal_live_reference(model, size):
    """The UNMODIFIED reference (ragged sizes: its reflection paddings are exercised) vs the restatement."""
    from oracle import reference_shims as rs
    system, args = rs.build_system(model=model, loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=1)
    seeded = bb.seeded_params(model, args.random_seed)
    init = {k: v.detach().clone() for k, v in system.net.named_parameters()}
    assert list(init) == list(seeded) and all(torch.equal(init[k], seeded[k]) for k in init)
    g = torch.Generator().manual_seed(7)
    frames = [torch.rand(1, 3, *size, generator=g) - 0.4 for _ in range(7)]
    ora = maml.OracleSystem(model, init, optimizer="SGD", num_steps=1)
    loss, preds, _, _ = ora.run_train_iter(frames, 0)
    losses, rpreds, _ = system.run_train_iter(frames, epoch=0, do_evaluation=False)
    assert float(losses["loss"]) == pytest.approx(float(loss), abs=1e-7 * max(1.0, abs(float(loss))))
    assert (rpreds[0] - preds[0]).abs().max().item() <= 1e-7 * max(1.0, preds[0].abs().max().item())
    post = {k: v.detach() for k, v in system.net.named_parameters()}
    assert max((post[k] - ora.params[k].detach()).abs().max().item() for k in post) <= 1e-9


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_oracle_equals_live_reference():
    from oracle import reference_shims as rs
    system, args = rs.build_system(model="sepconv", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=1)
    g = torch.Generator().manual_seed(7)
    frames = [torch.rand(1, 3, 40, 40, generator=g) for _ in range(7)]
    ora = maml.OracleSystem("sepconv", {k: v.detach().clone() for k, v in system.net.named_parameters()},
                            optimizer="SGD", num_steps=1)
    loss, preds, _, _ = ora.run_train_iter(frames, 0)
    losses, rpreds, _ = system.run_train_iter(frames, epoch=0, do_evaluation=False)
    assert float(losses["loss"]) == pytest.approx(float(loss), abs=1e-7)
    assert (rpreds[0] - preds[0]).abs().max().item() <= 1e-7


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_outer_optimizer_state_interchanges_with_reference_checkpoints():
    """the fused flat-buffer outer optimizer speaks the stock ``torch.optim.Adam.state_dict()``
    layout the reference checkpoints carry (utils.py:34-118): after one meta-iteration both hold the same per-parameter
    moments; a fresh system restored from the REFERENCE's model + optimizer state continues exactly like the
    reference; and the reference's Adam accepts the state this repo writes."""
    import copy
    import warnings
    from oracle import reference_shims as rs
    from oracle.ops_ref import RefOps
    from oracle.make_golden import synthetic_frames
    from meta_interpolation_b200 import backbone
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    from helpers import make_args

    over = dict(model="sepconv", loss="1*L1", optimizer="Adam", number_of_training_steps_per_iter=1, outer_lr=1e-3)
    ref, rargs = rs.build_system(batch_size=1, **over)
    ops = RefOps()
    saved = backbone._default_ops
    backbone.set_default_ops(ops)
    try:
        # (compat path on purpose: this test is about the optimizer state layout, and its moment tolerances are set
        # for the reference's own order of operations; the graph path's Adam inner rule is held against the compat
        # path in tests/test_host_logic.py)
        mine = SceneAdaptiveInterpolation(make_args(batch_size=1, fast_path=False, **over), ops=ops)
        frames = synthetic_frames(3, 1, 32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.run_train_iter(frames, epoch=0)
        mine.run_train_iter(frames, epoch=0)
        rsd, msd = ref.optimizer.state_dict(), mine.optimizer.state_dict()
        assert sorted(rsd["state"]) == sorted(msd["state"]) and len(msd["state"]) == len(list(mine.trainable_parameters()))
        for i, e in rsd["state"].items():
            assert int(float(e["step"])) == int(float(msd["state"][i]["step"])) == 1
            for key in ("exp_avg", "exp_avg_sq"):
                a, b = msd["state"][i][key], e[key]
                assert a.shape == b.shape
                assert (a - b).abs().max().item() <= 2e-5 * max(b.abs().max().item(), 1e-12), (i, key)

        # reference checkpoint -> this repo: model + optimizer state, then one more iteration on both
        fresh = SceneAdaptiveInterpolation(make_args(batch_size=1, fast_path=False, **over), ops=ops)
        fresh.load_state_dict(copy.deepcopy(ref.state_dict()))
        fresh.optimizer.load_state_dict(copy.deepcopy(rsd))
        frames2 = synthetic_frames(4, 1, 32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.run_train_iter(frames2, epoch=0)
        fresh.run_train_iter(frames2, epoch=0)
        rp, fp = dict(ref.net.named_parameters()), dict(fresh.net.named_parameters())
        for k in rp:
            d = (rp[k].detach() - fp[k].detach()).abs().max().item()
            # an Adam step is lr * m / sqrt(v): elements whose gradient is ~0 amplify rounding, so the bar is a small
            # fraction of the step size (lr = 1e-3), not of the parameter scale
            mean = (rp[k].detach() - fp[k].detach()).abs().mean().item()
            assert d <= 0.1 * over["outer_lr"] and mean <= 1e-3 * over["outer_lr"], (k, d, mean)

        # this repo -> the reference's stock Adam
        ref2, _ = rs.build_system(batch_size=1, **over)
        ref2.optimizer.load_state_dict(copy.deepcopy(mine.optimizer.state_dict()))
        st = ref2.optimizer.state_dict()["state"]
        assert len(st) == len(msd["state"]) and torch.equal(st[1]["exp_avg"], msd["state"][1]["exp_avg"])
    finally:
        backbone.set_default_ops(saved)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_checkpoint_files_interchange_with_the_reference(tmp_path, monkeypatch):
    """SURVEY 8f rank 3 (utils.py:34-118, meta_learning_system.py:154-170): a checkpoint written by the REFERENCE's
    ``save_checkpoint`` resumes this package's system (``--resume``; ``model_best.pth`` in val mode); a checkpoint
    written here is restored by the REFERENCE's ``load_checkpoint``; ``--pretrained_model`` overlays a backbone file
    with the reference's lossy matching rules (unknown keys and shape mismatches are skipped)."""
    import contextlib
    import io
    from oracle import reference_shims as rs
    from oracle.ops_ref import RefOps
    from meta_interpolation_b200 import backbone, utils as my_utils
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    from helpers import make_args

    over = dict(model="sepconv", loss="1*L1", optimizer="Adam", number_of_training_steps_per_iter=2, metasgd=True)
    ref, rargs = rs.build_system(batch_size=1, **over)
    import utils as ref_utils
    monkeypatch.chdir(tmp_path)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():                       # move every tensor off its init so equality means "loaded"
        for p in ref.parameters():
            p.add_(torch.rand(p.shape, generator=g) * 1e-2)
    ref_state = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    ref_utils.save_checkpoint({"epoch": 3, "arch": rargs, "state_dict": ref.state_dict(), "best_PSNR": 31.5}, False,
                              "expA")
    best = {k: v + 1.0 for k, v in ref_state.items()}
    ref_utils.save_checkpoint({"epoch": 7, "arch": rargs, "state_dict": best, "best_PSNR": 33.0}, True, "expB")

    ops = RefOps()
    saved = backbone._default_ops
    backbone.set_default_ops(ops)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            # reference file -> this package, through the constructor's --resume branch
            a1 = make_args(batch_size=1, load_checkpoint=True, exp_name="expA", **over)
            mine = SceneAdaptiveInterpolation(a1, ops=ops)
            assert a1.start_epoch == 3 and a1.resume_exp == "expA"
            assert list(mine.state_dict()) == list(ref_state)
            for k, v in mine.state_dict().items():
                assert torch.equal(v, ref_state[k]), k
            # val mode reads model_best.pth; resuming another experiment restarts the epoch counter
            a2 = make_args(batch_size=1, load_checkpoint=True, exp_name="fresh", resume_exp="expB", mode="val", **over)
            other = SceneAdaptiveInterpolation(a2, ops=ops)
            assert a2.start_epoch == 0
            for k, v in other.state_dict().items():
                assert torch.equal(v, best[k]), k

            # this package -> the reference's own loader
            with torch.no_grad():
                for p in mine.parameters():
                    p.mul_(1.5)
            my_utils.save_checkpoint({"epoch": 11, "arch": a1, "state_dict": mine.state_dict(), "best_PSNR": 30.0},
                                     True, "expC")
            assert sorted(os.listdir("checkpoint/expC")) == ["checkpoint.pth", "model_best.pth"]
            ref2, rargs2 = rs.build_system(batch_size=1, exp_name="expC", resume_exp=None, mode="train", **over)
            # (the reference calls torch.load(path) bare; torch >= 2.6 then refuses the pickled Namespace under 'arch',
            # for its own files too -- give its call the torch default it was written against)
            plain_load = torch.load
            monkeypatch.setattr(torch, "load", lambda f, *a, **k: plain_load(f, *a, **dict(k, weights_only=False)))
            ref_utils._original_load_checkpoint(rargs2, ref2, None)
            monkeypatch.setattr(torch, "load", plain_load)
            assert rargs2.start_epoch == 11
            want = mine.state_dict()
            for k, v in ref2.state_dict().items():
                assert torch.equal(v, want[k]), k

            # --pretrained_model: lossy overlay of a backbone file
            net_state = {k: v + 2.0 for k, v in mine.net.state_dict().items()}
            first, second = list(net_state)[:2]
            net_state[first] = torch.zeros(1, 2, 3)                  # wrong shape: skipped
            net_state["not.a.parameter"] = torch.zeros(4)            # unknown key: ignored
            torch.save({"state_dict": net_state}, "backbone.pth")
            a3 = make_args(batch_size=1, pretrained_model="backbone.pth", **over)
            third = SceneAdaptiveInterpolation(a3, ops=ops)
            fresh = SceneAdaptiveInterpolation(make_args(batch_size=1, **over), ops=ops)
            got, init = third.net.state_dict(), fresh.net.state_dict()
            assert torch.equal(got[first], init[first])
            assert torch.equal(got[second], net_state[second])
    finally:
        backbone.set_default_ops(saved)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_super_loss_equals_live_reference():
    """oracle/super_loss.py against the reference's SuperSloMoLoss (loss.py:246-274) on the auxiliary outputs of the
    SuperSloMo oracle, value and gradient; torchvision's vgg16() is handed the same seeded random weights on the
    reference side (the ImageNet weights cannot be downloaded here)."""
    from oracle import reference_shims as rs, super_loss as sl, backbones_flow as bf
    rs.install()
    import loss as ref_loss
    import torchvision.models as tv_models
    state = sl.seeded_vgg16_state(0)
    real = tv_models.vgg16

    def seeded(pretrained=False, **kw):
        m = real(weights=None)
        m.load_state_dict(state, strict=False)
        return m
    saved = ref_loss.models.vgg16
    ref_loss.models.vgg16 = seeded
    try:
        crit = ref_loss.SuperSloMoLoss()
    finally:
        ref_loss.models.vgg16 = saved
    params = {k: v.clone().requires_grad_(True) for k, v in bb.seeded_params("superslomo", 12345).items()}
    g = torch.Generator().manual_seed(1)
    i0, i1, hr = (torch.rand(2, 3, 64, 64, generator=g) - 0.4 for _ in range(3))
    out, aux = bf.superslomo_forward(i0, i1, params, params, full=True)
    a = crit(out, hr, I0=i0, I1=i1, **aux)
    b = sl.super_loss(out, hr, aux, i0, i1, state)
    assert abs(float(a) - float(b)) <= 1e-6 * abs(float(a))
    ga = torch.autograd.grad(a, list(params.values()), retain_graph=True)
    gb = torch.autograd.grad(b, list(params.values()))
    for k, x, y in zip(params, ga, gb):
        assert (x - y).abs().max().item() <= 1e-6 * max(1.0, x.abs().max().item()), k
