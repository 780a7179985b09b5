"""The staging kernel (mi_septuplet_prepare) on the GPU: bit-exact against the reference-generated golden vectors, the
oracle at the full Vimeo geometry, and end to end through the data provider into a meta-iteration."""
import random

import pytest
import torch

from helpers import make_args
from oracle.ops_ref import RefOps
from test_data_path import _args, run_case, staging_cases

pytestmark = pytest.mark.gpu


def test_staging_kernel_equals_reference_goldens(cuda_ops):
    for c in staging_cases():
        n0 = cuda_ops.launch_count()
        out = run_case(cuda_ops, c, "cuda")
        assert cuda_ops.launch_count() == n0 + 1                     # one launch per meta-batch
        assert torch.equal(out.cpu(), c["expect"]), (c["model"], c["mode"])


@pytest.mark.parametrize("model", ["sepconv", "superslomo", "voxelflow"])
def test_staging_kernel_full_size_equals_oracle(cuda_ops, model):
    """BASELINE geometry: 8 tasks of 7 decoded 256x448 frames, 256x256 random crops, random temporal flips."""
    from meta_interpolation_b200.data.vimeo_septuplet import VimeoSeptuplet
    g = torch.Generator().manual_seed(8)
    raw = torch.randint(0, 256, (8, 7, 256, 448, 3), generator=g, dtype=torch.uint8)
    case = {"raw": raw, "y0": [0] * 8, "x0": torch.randint(0, 193, (8,), generator=g).tolist(), "h": 256, "w": 256,
            "reversed": [bool(i % 3 == 0) for i in range(8)], "div255": model != "voxelflow"}
    case["mean"], case["std"] = VimeoSeptuplet.NORMALISATION.get(model, (None, None))
    want = run_case(RefOps(), case)
    got = run_case(cuda_ops, case, "cuda")
    assert torch.equal(got.cpu(), want)
    full = dict(case, x0=[0] * 8, h=256, w=448)                        # validation: whole frames
    assert torch.equal(run_case(cuda_ops, full, "cuda").cpu(), run_case(RefOps(), full))


def test_staging_rejects_bad_arguments(cuda_ops):
    from meta_interpolation_b200._lib import MiB200Error
    raw = torch.zeros(1, 7, 8, 8, 3, dtype=torch.uint8, device="cuda")
    z = torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(MiB200Error):
        cuda_ops.septuplet_prepare(raw, z, z, torch.zeros(1, dtype=torch.uint8, device="cuda"), 9, 8)


def test_provider_feeds_a_meta_iteration(cuda_ops, tmp_path):
    """PNG tree -> DataLoader (decode only, pinned uint8) -> one staging launch -> run_train_iter; the staged frames
    equal the oracle route on the same seed."""
    from oracle.make_golden_data import build_tree
    from meta_interpolation_b200.data import MetaLearningSystemDataLoader
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    build_tree(str(tmp_path), n_train=2, n_test=1, h=72, w=80)
    res = {}
    for name, ops in (("cuda", cuda_ops), ("ref", RefOps())):
        provider = MetaLearningSystemDataLoader(_args(tmp_path), ops=ops)
        provider.dataset.crop_size = 64
        torch.manual_seed(1); random.seed(1)
        res[name] = list(provider.get_train_batches())
    (frames, meta), (want, _) = res["cuda"][0], res["ref"][0]
    assert len(res["cuda"]) == 1 and len(frames) == 7 and frames[0].is_cuda and frames[0].shape == (2, 3, 64, 64)
    for a, b in zip(frames, want):
        assert torch.equal(a.cpu(), b)
    system = SceneAdaptiveInterpolation(make_args(cuda=True, batch_size=2), ops=cuda_ops)
    losses, preds, _ = system.run_train_iter(frames, epoch=0)
    assert torch.isfinite(losses["loss"]).item()
    assert tuple(preds[0].shape[-3:]) == (3, 64, 64)
