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
