"""Task sharding + meta-gradient all-reduce: a 2-rank gloo run on CPU must reproduce the 1-rank run
of the same meta-batch (sum of per-rank mean gradients scaled by 1/R == global mean)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _extra_state(system):
    """gamma_mult / attenuator parameters of an L2F system (their outer gradients cross the all-reduce too)."""
    return {k: v.detach().clone() for k, v in system.state_dict().items()
            if k == "gamma_mult" or k.startswith("attenuator")}


def _prepare(system, l2f):
    if l2f:      # away from the init (gamma_mult = 0), where the attenuator gradients vanish
        with torch.no_grad():
            system.gamma_mult.fill_(0.3)
        system.optimizer.param_groups[0]["lr"] = 1.0


def _worker(rank, world, port, out_dir, l2f=False, batch=2):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import make_args
    from oracle.ops_ref import RefOps
    from meta_interpolation_b200 import backbone
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    ops = RefOps()
    backbone.set_default_ops(ops)
    for fast in (True, False):          # both execution paths in one rendezvous (process start-up dominates)
        args = make_args(batch_size=batch, number_of_training_steps_per_iter=1, fast_path=fast, attenuate=l2f)
        system = SceneAdaptiveInterpolation(args, ops=ops)
        _prepare(system, l2f)
        g = torch.Generator().manual_seed(11)
        frames = [torch.rand(batch, 3, 32, 32, generator=g) for _ in range(7)]
        losses, preds, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
        mine = [i for i, p in enumerate(preds) if torch.is_tensor(p)]
        torch.save({"flat": system.net.arena.flat.clone(), "tasks": mine, "loss": float(losses["loss"]),
                    "total": float(losses["total"]), "psnr": float(metrics["psnr"].avg),
                    "extra": _extra_state(system)}, os.path.join(out_dir, "rank%d_fast%d.pt" % (rank, int(fast))))
    dist.destroy_process_group()


def _single(fast, l2f=False, batch=2, full=False):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_args
    from oracle.ops_ref import RefOps
    from meta_interpolation_b200 import backbone
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    ops = RefOps()
    saved = backbone._default_ops
    backbone.set_default_ops(ops)
    try:
        args = make_args(batch_size=batch, number_of_training_steps_per_iter=1, fast_path=fast, attenuate=l2f)
        system = SceneAdaptiveInterpolation(args, ops=ops)
        _prepare(system, l2f)
        g = torch.Generator().manual_seed(11)
        frames = [torch.rand(batch, 3, 32, 32, generator=g) for _ in range(7)]
        losses, _, metrics = system.run_train_iter(frames, epoch=0, do_evaluation=True)
        if full:
            return (system.net.arena.flat.clone(), float(losses["loss"]), float(losses["total"]),
                    float(metrics["psnr"].avg))
        if l2f:
            return system.net.arena.flat.clone(), float(losses["loss"]), _extra_state(system)
        return system.net.arena.flat.clone(), float(losses["loss"])
    finally:
        backbone.set_default_ops(saved)


def _load(tmp_path, rank, fast):
    return torch.load(os.path.join(str(tmp_path), "rank%d_fast%d.pt" % (rank, int(fast))))


def test_two_ranks_equal_one_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for fast in (True, False):
        r0, r1 = _load(tmp_path, 0, fast), _load(tmp_path, 1, fast)
        assert r0["tasks"] == [0] and r1["tasks"] == [1]           # each rank adapted its own shard
        assert torch.equal(r0["flat"], r1["flat"])                   # identical outer step on every rank
        single, loss = _single(fast)
        # outer SGD step: theta - lr * mean-gradient; the two summation orders agree to fp32 rounding
        assert (r0["flat"] - single).abs().max().item() <= 1e-9, fast
        # the logged loss is all-reduced: every rank reports the meta-batch's value, not its shard's
        assert abs(r0["loss"] - loss) <= 1e-6 and abs(r1["loss"] - loss) <= 1e-6, fast


@pytest.mark.parametrize("batch,shards", [(3, ([0], [1, 2])), (1, ([], [0]))])
def test_meta_batch_not_divisible_by_world_size(tmp_path, batch, shards):
    """The short last batch of an epoch (DataLoader drop_last=False): tasks are split raggedly, every outer gradient
    is scaled by 1/B, so the SUM all-reduce is still the mean over the meta-batch; a rank with no task only takes
    part in the collective and the step."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), False, batch), nprocs=2, join=True)
    for fast in (True, False):
        r0, r1 = _load(tmp_path, 0, fast), _load(tmp_path, 1, fast)
        assert (r0["tasks"], r1["tasks"]) == tuple(list(s) for s in shards)
        assert torch.equal(r0["flat"], r1["flat"])
        single, loss, total, psnr = _single(fast, batch=batch, full=True)
        assert (r0["flat"] - single).abs().max().item() <= 1e-8, fast
        for r in (r0, r1):
            assert abs(r["loss"] - loss) <= 1e-6 and abs(r["total"] - total) <= 1e-6, fast
            assert abs(r["psnr"] - psnr) <= 1e-4, fast


def test_two_ranks_equal_one_rank_l2f(tmp_path):
    """BASELINE configs[3] shape of the problem (L2F sharded over ranks): the attenuator / gamma_mult outer gradients
    are gathered into their flat group BEFORE the all-reduce on the graph path, so two ranks reproduce one."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), True), nprocs=2, join=True)
    for fast in (True, False):
        r0, r1 = _load(tmp_path, 0, fast), _load(tmp_path, 1, fast)
        assert torch.equal(r0["flat"], r1["flat"])
        single, loss, extra = _single(fast, True)
        scale = max(1.0, single.abs().max().item())
        assert (r0["flat"] - single).abs().max().item() <= 1e-5 * scale, fast
        assert abs(r0["loss"] - loss) <= 1e-6 and abs(r1["loss"] - loss) <= 1e-6, fast
        for k, v in extra.items():
            assert torch.equal(r0["extra"][k], r1["extra"][k]), k
            d = (r0["extra"][k] - v).abs().max().item()
            assert d <= 1e-4 * max(v.abs().max().item(), 1e-6), (k, d, fast)
