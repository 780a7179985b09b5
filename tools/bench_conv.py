"""GPU micro-benchmark: per-layer TFLOP/s of the conv engines at the SepConv layer shapes (N=2 support batch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402
from meta_interpolation_b200.ops import ENGINE_TC, WG_STORE, WgradSpec, pad4  # noqa: E402

SHAPES = [(2, 384, 512, 32, 32), (2, 384, 512, 51, 51), (2, 192, 256, 64, 64), (2, 192, 256, 64, 51),
          (2, 96, 128, 128, 128), (2, 48, 64, 256, 256), (2, 24, 32, 512, 512), (2, 12, 16, 512, 512),
          # region-of-interest Subnet shapes and the channel-changing mid layers
          (2, 258, 450, 51, 51), (2, 137, 233, 64, 64), (2, 96, 128, 64, 128), (2, 96, 128, 128, 64),
          (2, 48, 64, 128, 256), (2, 48, 64, 256, 128), (2, 48, 64, 128, 128), (2, 24, 32, 256, 512),
          (2, 24, 32, 512, 256), (2, 24, 32, 256, 256)]


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
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("env HALO=%s HALO_STREAM=%s WGRAD_KX=%s" % (os.environ.get("MI_B200_HALO"),
                                                      os.environ.get("MI_B200_HALO_STREAM"),
                                                      os.environ.get("MI_B200_WGRAD_KX")))
    for (n, h, w, cin, cout) in SHAPES:
        x = ops.empty_act(n, h, w, cin); x.copy_(torch.rand(n, h, w, cin, device="cuda") - 0.5)
        dy = ops.empty_act(n, h, w, cout); dy.copy_(torch.rand(n, h, w, cout, device="cuda") - 0.5)
        wt = ops.empty_weight(cout, cin, 3); wt.copy_(torch.rand(cout, 3, 3, cin, device="cuda") - 0.5)
        b = torch.rand(cout, device="cuda")
        y = ops.empty_act(n, h, w, cout)
        gw, gb = ops.empty_weight(cout, cin, 3), torch.zeros(cout, device="cuda")
        flops = 2.0 * n * h * w * cin * cout * 9
        byts = 4.0 * (n * h * w * (cin + cout) + cout * 9 * cin)
        line = "%-24s" % str((n, h, w, cin, cout))
        if which in ("all", "fprop"):
            ms = timeit(lambda: ops.conv_fprop(x, wt, b, 1, 0.0, out=y, engine=ENGINE_TC))
            line += " fprop %7.1f us %6.1f TF/s %6.0f GB/s |" % (ms * 1e3, flops / ms / 1e9, byts / ms / 1e6)
        if which in ("all", "wgrad"):
            spec = WgradSpec(WG_STORE, grad_w=gw, grad_b=gb)
            ms = timeit(lambda: ops.conv_wgrad(x, dy, 3, pad4(cin), spec, engine=ENGINE_TC))
            line += " wgrad %7.1f us %6.1f TF/s" % (ms * 1e3, flops / ms / 1e9)
        print(line)


if __name__ == "__main__":
    main()
