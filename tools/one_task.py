"""One 256x448 SepConv task (K=5 inner steps + query + outer step) with eager launches, bracketed by
cudaProfilerStart/Stop: the target of the ncu launch list committed under profiles/.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/one_task.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation  # noqa: E402

args = bench.make_args(1)
args.cuda_graphs = False
system = SceneAdaptiveInterpolation(args)
frames = [f.cuda() for f in bench.synthetic_septuplets(1, 7)]
system.run_train_iter(frames, epoch=0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
system.run_train_iter(frames, epoch=0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
