"""GPU micro-benchmark of the input staging kernel at the bench geometry (8 tasks x 7 frames x 256x448):
device-resident uint8 frames -> float NCHW frames, and the same including the pinned host-to-device copy, against
shipping ready-made float frames (what the reference's DataLoader does)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ops = default_ops()
    tasks, frames, h, w = 8, 7, 256, 448
    host = torch.randint(0, 256, (tasks, frames, h, w, 3), dtype=torch.uint8).pin_memory()
    dev = host.cuda()
    z = torch.zeros(tasks, dtype=torch.int32, device="cuda")
    rev = torch.zeros(tasks, dtype=torch.uint8, device="cuda")
    floats = torch.rand(frames, tasks, 3, h, w).pin_memory()
    sink = torch.empty_like(floats, device="cuda")
    k = timeit(lambda: ops.septuplet_prepare(dev, z, z, rev, h, w))
    e2e = timeit(lambda: ops.septuplet_prepare(host.cuda(non_blocking=True), z, z, rev, h, w))
    ref = timeit(lambda: sink.copy_(floats, non_blocking=True))
    px = tasks * frames * h * w
    print(json.dumps({"staging_kernel_us": round(k * 1e3, 1), "kernel_GBps": round(px * 15 / k / 1e6, 1),
                      "uint8_h2d_plus_kernel_us": round(e2e * 1e3, 1), "float_h2d_us": round(ref * 1e3, 1),
                      "bytes_u8": px * 3, "bytes_f32": px * 12}))


if __name__ == "__main__":
    main()
