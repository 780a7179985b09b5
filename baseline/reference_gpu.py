"""Run the UNMODIFIED reference (staged copy under baseline/_ref/, git-ignored) on ONE GPU: the denominator of
north_star's ">= 10x the reference's single-GPU PyTorch inner-loop throughput" and a live fp32 ground truth for
full-size parity.  Reference arm / test infrastructure; none of this repo's kernels or engine run here.

    CUDA_VISIBLE_DEVICES=0 python baseline/reference_gpu.py --model sepconv --batch 8 --steps 5 --warmup 3 \
        [--no-tf32] [--dump out.pt]

Shims (no reference file is modified): `cupy` -> NVRTC + cuLaunchKernel stand-in so the
reference's own separable-convolution kernel strings run (baseline/cupy_nvrtc.py); `ReduceLROnPlateau(verbose=)`
accepted; `utils.load_checkpoint` no-op with `args.resume=True` (no pretrained files offline).  Exactly one GPU
must be visible (reference meta_learning_system.py:285-289 takes a DataParallel branch otherwise).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MI_REFERENCE_ROOT", os.path.join(HERE, "_ref"))

CONFIGS = {
    # BASELINE.json configs[1..4]: reference flags, frame size
    "sepconv": (dict(model="sepconv", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=5), (256, 448)),
    "superslomo": (dict(model="superslomo", loss="1*L1", optimizer="SGD", metasgd=True,
                        number_of_training_steps_per_iter=5), (256, 448)),
    "cain": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True, number_of_training_steps_per_iter=3),
             (512, 512)),
    "rrin": (dict(model="rrin", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=5,
                  learnable_per_layer_per_step_inner_loop_learning_rate=True, use_multi_step_loss_optimization=True),
             (256, 448)),
}


def available():
    return os.path.isfile(os.path.join(REF, "meta_learning_system.py"))


def install():
    import torch
    sys.path.insert(0, ROOT)
    from baseline import cupy_nvrtc
    cupy_nvrtc.install()
    import torch.optim.lr_scheduler as lrs
    if not getattr(lrs.ReduceLROnPlateau, "_mi_shim", False):
        _orig = lrs.ReduceLROnPlateau

        class ReduceLROnPlateau(_orig):
            _mi_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        lrs.ReduceLROnPlateau = ReduceLROnPlateau
        torch.optim.lr_scheduler.ReduceLROnPlateau = ReduceLROnPlateau
    if REF not in sys.path:
        sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)      # sepconv/model.py adds a cwd-relative sys.path entry
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import utils as ref_utils
            ref_utils.load_checkpoint = lambda *a, **k: None
            import sepconv.model  # noqa: F401
    finally:
        os.chdir(cwd)


def build_system(over, batch):
    import torch
    assert torch.cuda.device_count() == 1, "run with CUDA_VISIBLE_DEVICES=<one GPU>"
    install()
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        import config as ref_config
        args, _ = ref_config.get_args()
    finally:
        sys.argv = argv
    args = type(args)(**vars(args))
    args.cuda, args.num_gpu, args.resume, args.batch_size = True, 1, True, batch
    for k, v in over.items():
        setattr(args, k, v)
    import warnings
    warnings.simplefilter("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        import meta_learning_system as ref_mls
        system = ref_mls.SceneAdaptiveInterpolation(args)
    return system, args


def frames_for(model, batch, seed, hw):
    """bench.synthetic_septuplets + the dataset's per-model normalisation (data/vimeo_septuplet.py:31-40,73-76)."""
    import torch
    import bench
    frames = bench.synthetic_septuplets(batch, seed, *hw)
    if model == "superslomo":
        mean = torch.tensor([0.429, 0.431, 0.397]).view(1, 3, 1, 1)
        frames = [f - mean for f in frames]
    elif model == "voxelflow":
        frames = [(f * 255 - 127.5) / 127.5 for f in frames]
    return frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sepconv", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-tf32", dest="tf32", action="store_false")
    ap.add_argument("--weight-gain", type=float, default=None, help="scale every 4-D weight of the seeded init")
    ap.add_argument("--dump", default=None, help="write loss / preds / PSNR of the FIRST iteration (parity fixture)")
    ap.add_argument("--seed", type=int, default=100)
    a = ap.parse_args()
    import torch
    if not available():
        print(json.dumps({"impl": "reference_gpu", "unavailable": "baseline/_ref is not staged"}))
        return
    torch.backends.cudnn.allow_tf32 = a.tf32
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    over, hw = CONFIGS[a.model]
    system, args = build_system(over, a.batch)
    if a.weight_gain is not None:
        with torch.no_grad():
            for p in system.net.parameters():
                if p.dim() == 4:
                    p.mul_(a.weight_gain)
    sets = [[f.cuda() for f in frames_for(a.model, a.batch, a.seed + s, hw)] for s in range(2)]
    first = None
    times = []
    for i in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        losses, preds, metrics = system.run_train_iter(sets[i % 2], epoch=0, do_evaluation=(i == 0 and a.dump is not None))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i == 0:
            first = (float(losses["loss"]), torch.cat([p.detach().reshape(1, *p.shape[-3:]) for p in preds]).cpu(),
                     float(metrics["psnr"].avg) if a.dump else None)
        if i >= a.warmup:
            times.append(dt)
    per_iter = sum(times) / len(times)
    line = {"impl": "reference_gpu", "model": a.model, "tasks_per_s": round(a.batch / per_iter, 4),
            "ms_per_iter": round(per_iter * 1e3, 2), "batch": a.batch, "steps": a.steps, "warmup": a.warmup,
            "allow_tf32": bool(a.tf32), "loss_first_iter": first[0], "gpu": torch.cuda.get_device_name(0),
            "torch": torch.__version__,
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    if a.dump:
        torch.save({"loss": first[0], "preds": first[1], "psnr": first[2], "model": a.model, "batch": a.batch,
                    "seed": a.seed, "tf32": bool(a.tf32), "weight_gain": a.weight_gain}, a.dump)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
