"""One task of another BASELINE configuration with eager launches between cudaProfilerStart/Stop (ncu target).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/one_task_model.py superslomo
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bench_backbones import CONFIGS, normalise  # noqa: E402
from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation  # noqa: E402

name = sys.argv[1]
over, (h, w), _, _ = CONFIGS[name]
a = bench.make_args(1)
for k, v in over.items():
    setattr(a, k, v)
a.cuda_graphs = False
system = SceneAdaptiveInterpolation(a)
frames = [f.cuda() for f in normalise(bench.synthetic_septuplets(1, 7, h, w), name)]
system.run_train_iter(frames, epoch=0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
system.run_train_iter(frames, epoch=0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
