"""GPU micro-benchmark of the x2 bilinear resampling kernels (csrc/elementwise.cu, upsample2_*): achieved bytes/s against
the measured HBM copy bandwidth.  Algorithmic traffic: 4 * C * (input + output) bytes (every value read
once and written once).  Geometries: the SepConv Subnet's windowed upsample (137x233 of a 192x256 grid -> 258x450 of
384x512, 64 -> the layer before the 51-channel head) and the decoder's plain upsamples.

    python tools/bench_upsample.py            # MI_B200_UPSAMPLE_STRIP=0: per-pixel forms; MI_B200_UPSAMPLE_ROWS=r: fixed strip
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402
from bench_warp import timeit  # noqa: E402


def main():
    ops = default_ops()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(root, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    tag = {"strip": os.environ.get("MI_B200_UPSAMPLE_STRIP", "1"), "rows": os.environ.get("MI_B200_UPSAMPLE_ROWS", "auto")}
    n = 2
    cases = [("window", 64, (137, 233), (192, 256), (27, 11), (258, 450), (63, 31)),
             ("window", 51, (137, 233), (192, 256), (27, 11), (258, 450), (63, 31)),
             ("plain", 64, (96, 128), None, None, None, None),
             ("plain", 128, (48, 64), None, None, None, None),
             ("plain", 512, (12, 16), None, None, None, None)]
    for kind, c, (h, w), full, lo, hi_hw, hi in cases:
        x = ops.empty_act(n, h, w, c); x.copy_(torch.rand(n, h, w, c, device="cuda"))
        if kind == "window":
            fwd = lambda: ops.upsample_window_fwd(x, True, full, lo, hi, hi_hw)
            y = fwd()
            dy = ops.empty_act(*y.shape); dy.copy_(torch.rand(*y.shape, device="cuda"))
            dx = ops.empty_act(n, h, w, c)
            bwd = lambda: ops.upsample_window_bwd(dy, dx, True, False, full, lo, hi, mask_y=x, mask_act=1, mask_slope=0.0)
        else:
            y = ops.empty_act(n, 2 * h, 2 * w, c)
            fwd = lambda: ops.upsample_fwd(x, True, out=y)
            dy = ops.empty_act(*y.shape); dy.copy_(torch.rand(*y.shape, device="cuda"))
            dx = ops.empty_act(n, h, w, c)
            bwd = lambda: ops.upsample_bwd(dy, dx, True, False)
        ld = x.stride(-2)
        nbytes = 4.0 * ld * (x.shape[0] * x.shape[1] * x.shape[2] + y.shape[0] * y.shape[1] * y.shape[2])
        for name, fn, extra in (("upsample2_fwd", fwd, 0.0), ("upsample2_bwd", bwd, 4.0 * ld * n * h * w if kind == "window" else 0.0)):
            ms = timeit(fn, flush)
            print(json.dumps(dict(tag, kernel=name, kind=kind, c=c, lo=[h, w], hi=list(y.shape[1:3]), us=round(ms * 1e3, 2),
                                  algorithmic_MB=round((nbytes + extra) / 1e6, 1),
                                  algorithmic_GBps=round((nbytes + extra) / ms / 1e6, 1),
                                  frac_of_hbm=round((nbytes + extra) / ms / 1e6 / hbm, 3), hbm_peak_GBps=hbm)))


if __name__ == "__main__":
    main()
