"""GPU micro-benchmark of the adaptive separable convolution kernels at the BASELINE geometry (N=2, 256x448 window of
the 258x450 region-of-interest filter maps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
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
    n, h, w, gh, gw = 2, 256, 448, 258, 450
    frame = torch.rand(n, 3, h, w, device="cuda")
    v, hh = ops.empty_act(n, gh, gw, 51), ops.empty_act(n, gh, gw, 51)
    v.copy_(torch.rand(n, gh, gw, 51, device="cuda")); hh.copy_(torch.rand(n, gh, gw, 51, device="cuda"))
    go = torch.rand(n, 3, h, w, device="cuda")
    gv, ghh = ops.zeros_act(n, gh, gw, 51), ops.zeros_act(n, gh, gw, 51)
    px = n * h * w
    for planar in (False, True):
        ws = ops.sepconv_planar(n, h, w, 51) if planar else None
        ws2 = ops.sepconv_planar(n, h, w, 51) if planar else None
        tag = "planar filters, 4 px/thread (transposes included)" if planar else "NHWC filters, 2 px/thread"
        ms = timeit(lambda: ops.sepconv_fwd(frame, v, hh, h, w, 1, 1, -25, -25, planar=ws))
        print("%-52s fwd %7.1f us  %5.1f TFLOP/s fp32" % (tag, ms * 1e3, 2.0 * px * (3 * 51 * 51 + 3 * 51) / ms / 1e9))
        ms = timeit(lambda: ops.sepconv_bwd(frame, v, hh, go, gv, ghh, 1, 1, -25, -25, planar=ws, planar_valid=planar,
                                            planar_grad=ws2))
        print("%-52s bwd %7.1f us  %5.1f TFLOP/s fp32" % (tag, ms * 1e3,
                                                         2.0 * px * (2 * 3 * 51 * 51 + 2 * 3 * 51) / ms / 1e9))


if __name__ == "__main__":
    main()
