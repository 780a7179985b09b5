"""Input staging: the dataset / data provider mirror of the reference's data/vimeo_septuplet.py and
the oracle of the staging kernel, against golden vectors produced by the unmodified reference dataset and -- in the
build container -- against the reference dataset itself on a miniature Vimeo tree."""
import os
import random

import pytest
import torch

from helpers import GOLDEN_DIR
from oracle.ops_ref import RefOps


def staging_cases():
    return torch.load(os.path.join(GOLDEN_DIR, "septuplet_staging.pt"), weights_only=False)


def run_case(ops, case, device="cpu"):
    dev = torch.device(device)
    return ops.septuplet_prepare(case["raw"].to(dev), torch.tensor(case["y0"], dtype=torch.int32, device=dev),
                                 torch.tensor(case["x0"], dtype=torch.int32, device=dev),
                                 torch.tensor(case["reversed"], device=dev).to(torch.uint8), case["h"], case["w"],
                                 bgr=True, div255=case["div255"], mean=case["mean"], std=case["std"])


def test_staging_oracle_equals_reference_goldens():
    cases = staging_cases()
    assert {(c["model"], c["mode"]) for c in cases} == {(m, s) for m in ("sepconv", "superslomo", "voxelflow")
                                                        for s in ("train", "val")}
    assert any(any(c["reversed"]) for c in cases) and any(not all(c["reversed"]) for c in cases)
    for c in cases:
        out = run_case(RefOps(), c)
        assert out.shape == c["expect"].shape
        assert torch.equal(out, c["expect"]), (c["model"], c["mode"])     # bit-exact: byte -> float arithmetic


def _args(tmp, **kw):
    import argparse
    d = dict(data_root=str(tmp), batch_size=2, val_batch_size=1, test_batch_size=1, mode="train", model="sepconv",
             dataset="vimeo90k", num_workers=0, num_gpu=0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_to_device_checks_windows_and_sizes(tmp_path):
    from oracle.make_golden_data import build_tree
    from meta_interpolation_b200.data.vimeo_septuplet import VimeoSeptuplet
    build_tree(str(tmp_path), n_train=2, n_test=1, h=20, w=24)
    ds = VimeoSeptuplet(_args(tmp_path), ops=RefOps())
    assert len(ds) == 2 and ds.crop_size == 256
    staged, meta = ds[0]
    assert staged["raw"].shape == (7, 20, 24, 3) and (staged["h"], staged["w"]) == (20, 24)   # smaller than the crop
    assert len(meta["imgpaths"]) == 7
    batch = {"raw": staged["raw"][None], "y0": torch.tensor([0]), "x0": torch.tensor([0]), "h": torch.tensor([20]),
             "w": torch.tensor([24]), "reversed": torch.tensor([False])}
    frames = ds.to_device(batch)
    assert len(frames) == 7 and frames[0].shape == (1, 3, 20, 24)
    with pytest.raises(ValueError):
        ds.to_device(dict(batch, y0=torch.tensor([1])))
    two = {k: torch.cat([v, v]) for k, v in batch.items()}
    two["h"] = torch.tensor([20, 16])
    with pytest.raises(ValueError):
        ds.to_device(two)
    ds.switch_set("val")
    assert len(ds) == 1


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("model", ["sepconv", "superslomo", "voxelflow"])
def test_dataset_and_provider_equal_live_reference(tmp_path, model):
    """Same list files, same ``random`` seed -> the reference's DataLoader batches (float frames it built on the CPU)
    equal the frames this package stages (decode only + one staging call), image paths included; train mode draws
    crops and temporal flips, val mode yields full frames."""
    import contextlib
    import io
    from oracle import reference_shims as rs
    from oracle.make_golden_data import build_tree
    from meta_interpolation_b200.data import MetaLearningSystemDataLoader
    rs.install()
    from data import MetaLearningSystemDataLoader as RefProvider
    build_tree(str(tmp_path), n_train=5, n_test=2, h=36, w=44)
    args = rs.make_args(model=model, mode="train", data_root=str(tmp_path), batch_size=2, num_workers=0)
    with contextlib.redirect_stdout(io.StringIO()):
        ref, mine = RefProvider(args), MetaLearningSystemDataLoader(args, ops=RefOps())
    ref.dataset.crop_size = mine.dataset.crop_size = 24
    for getter in ("get_train_batches", "get_val_batches"):
        torch.manual_seed(3); random.seed(3)
        want = list(getattr(ref, getter)())
        torch.manual_seed(3); random.seed(3)
        got = list(getattr(mine, getter)())
        assert len(want) == len(got) == (3 if getter == "get_train_batches" else 2)
        flips = 0
        for (wf, wm), (gf, gm) in zip(want, got):
            assert len(gf) == 7
            for a, b in zip(wf, gf):
                assert a.shape == b.shape and torch.equal(a, b)
            assert [list(p) for p in wm["imgpaths"]] == [list(p) for p in gm["imgpaths"]]
            flips += sum("im7.png" in p for p in gm["imgpaths"][0])
        if getter == "get_train_batches":
            assert wf[0].shape[-2:] == (24, 24) and 0 < flips < 5          # both orders occurred
    assert mine.total_train_iters_produced == ref.total_train_iters_produced


def test_provider_with_worker_processes(tmp_path):
    """Decode-only workers in separate processes (the reference's default is --num_workers 5): every staged batch
    equals the oracle staging of the raw frames and augmentation decisions the workers returned."""
    from torch.utils.data import DataLoader
    from oracle.make_golden_data import build_tree
    from meta_interpolation_b200.data import MetaLearningSystemDataLoader
    build_tree(str(tmp_path), n_train=4, n_test=1, h=30, w=34)
    provider = MetaLearningSystemDataLoader(_args(tmp_path, num_workers=2), ops=RefOps())
    provider.dataset.crop_size = 16
    provider.dataset.switch_set("train")
    seen = 0
    for staged, meta in DataLoader(provider.dataset, batch_size=2, shuffle=False, num_workers=2):
        frames = provider.dataset.to_device(staged)
        assert len(frames) == 7 and frames[0].shape == (2, 3, 16, 16)
        for b in range(2):
            y0, x0 = int(staged["y0"][b]), int(staged["x0"][b])
            order = range(6, -1, -1) if bool(staged["reversed"][b]) else range(7)
            for f, fs in enumerate(order):
                want = staged["raw"][b, fs, y0:y0 + 16, x0:x0 + 16][:, :, [2, 1, 0]].permute(2, 0, 1).float() / 255
                assert torch.equal(frames[f][b], want)
            assert ("im7.png" in meta["imgpaths"][0][b]) == bool(staged["reversed"][b])
        seen += 1
    assert seen == 2
    batches = list(provider.get_train_batches(total_batches=1))
    assert len(batches) == 1 and batches[0][0][0].shape == (2, 3, 16, 16)
