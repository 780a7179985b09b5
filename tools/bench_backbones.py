"""Tasks/s of `run_train_iter` for the other BASELINE configurations on ONE GPU (C1, C3, C4, C5 at their
per-GPU meta-batch), device-resident synthetic septuplets, CUDA-event timing, 3 warm-up + 3 timed iterations.

    python tools/bench_backbones.py [model ...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation  # noqa: E402

CONFIGS = {
    # name: (args overrides, (H, W), tasks per GPU, note)
    "voxelflow": (dict(model="voxelflow", loss="1*MSE", optimizer="SGD", number_of_training_steps_per_iter=1),
                  (128, 128), 1, "C1: voxelflow K=1 128x128 batch 1 (the reference's CPU plumbing case, here on the GPU)"),
    "superslomo": (dict(model="superslomo", loss="1*L1", optimizer="SGD", metasgd=True,
                        number_of_training_steps_per_iter=5), (256, 448), 4,
                   "C3: superslomo Meta-SGD K=5 256x448, 4 tasks per GPU (32 over 8)"),
    "cain": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True, number_of_training_steps_per_iter=3),
             (512, 512), 4, "C4: cain L2F K=3 512x512, 4 tasks per GPU (16 over 4)"),
    "rrin": (dict(model="rrin", loss="1*L1", optimizer="SGD", number_of_training_steps_per_iter=5,
                  learnable_per_layer_per_step_inner_loop_learning_rate=True, use_multi_step_loss_optimization=True,
                  multi_step_loss_num_epochs=1), (256, 448), 8,
             "C5: rrin MAML++ (MSL + learnable per-step lr) K=5 256x448, 8 tasks per GPU (64 over 8)"),
    # the operating point of the authors' scripts/run_sepconv.sh: Meta-SGD with the Adamax inner rule, K=3, batch 3
    "sepconv_adamax": (dict(model="sepconv", loss="1*L1", optimizer="Adamax", metasgd=True, inner_lr=1e-5,
                            outer_lr=1e-5, number_of_training_steps_per_iter=3), (256, 448), 3,
                       "scripts/run_sepconv.sh: sepconv Meta-SGD + Adamax K=3 256x448 batch 3, graph path"),
    "sepconv_adamax_compat": (dict(model="sepconv", loss="1*L1", optimizer="Adamax", metasgd=True, inner_lr=1e-5,
                                   outer_lr=1e-5, number_of_training_steps_per_iter=3, fast_path=False), (256, 448), 3,
                              "scripts/run_sepconv.sh: same, compat (autograd) path"),
}


def normalise(frames, model):
    if model == "superslomo":
        mean = torch.tensor([0.429, 0.431, 0.397]).view(1, 3, 1, 1)
        return [f - mean for f in frames]
    if model == "voxelflow":
        return [(f * 255 - 127.5) / 127.5 for f in frames]
    return frames


def main():
    names = sys.argv[1:] or list(CONFIGS)
    torch.cuda.set_device(0)
    for name in names:
        over, (h, w), batch, note = CONFIGS[name]
        a = bench.make_args(batch)
        for k, v in over.items():
            setattr(a, k, v)
        a.number_of_evaluation_steps_per_iter = a.number_of_training_steps_per_iter
        system = SceneAdaptiveInterpolation(a)
        if name == "cain":      # default init explodes through 125 stacked convs: seeded init x0.4
            with torch.no_grad():
                for p in system.net.parameters():
                    if p.dim() == 4:
                        p.mul_(0.4)
        frames = [f.cuda() for f in normalise(bench.synthetic_septuplets(batch, 321, h, w), name)]
        n0 = system.ops.launch_count()
        for _ in range(3):
            losses, _, _ = system.run_train_iter(frames, epoch=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        e0.record()
        for _ in range(steps):
            losses, _, _ = system.run_train_iter(frames, epoch=0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"config": note, "tasks_per_s": round(batch / (ms / 1e3), 3), "ms_per_meta_batch": round(ms, 2),
                          "tasks_per_gpu": batch, "fast_path": bool(system.fast_path_supported()),
                          "loss": round(float(losses["loss"]), 6),
                          "launches_per_task": int((system.ops.launch_count() - n0) / (6 * batch))}))
        del system
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
