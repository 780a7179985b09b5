"""GPU micro-benchmark of the bilinear backward-warp kernels (csrc/warp.cu): achieved bytes/s against the measured
HBM copy bandwidth.  SURVEY 8(d) algorithmic traffic: 4 * (3 + 2 + 3) bytes per pixel forward (image, flow, result),
4 * (3 + 2 + 3 + 2) backward; the buffers are 16-byte NHWC pixels, so what actually moves is 48 / 64 bytes per pixel.

    python tools/bench_warp.py [images]      # default 2 (the in-task launch) and 32 (a launch large enough to time)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402


def timeit(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()          # 256 MB > L2: the next launch reads from HBM
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ops = default_ops()
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "MEASURED_PEAKS.json"))) if os.path.isfile("MEASURED_PEAKS.json") else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    h, w = 256, 448
    for n in [int(a) for a in sys.argv[1:]] or [2, 32]:
        g = torch.Generator(device="cuda").manual_seed(0)
        img = ops.empty_act(n, h, w, 3); img.copy_(torch.rand(n, h, w, 3, device="cuda", generator=g))
        flow = ops.empty_act(n, h, w, 2)
        base = torch.nn.functional.interpolate(torch.randn(n, 2, h // 16, w // 16, device="cuda", generator=g) * 4,
                                               size=(h, w), mode="bilinear")
        flow.copy_(base.permute(0, 2, 3, 1))
        go = ops.empty_act(n, h, w, 3); go.copy_(torch.rand(n, h, w, 3, device="cuda", generator=g))
        gflow = ops.empty_act(n, h, w, 2)
        out = ops.empty_act(n, h, w, 3)
        px = n * h * w
        for variant in (0, 1):
            ms = timeit(lambda: ops.warp_fwd(img, flow, variant, out=out), flush)
            print(json.dumps({"kernel": "warp_fwd", "variant": variant, "images": n, "us": round(ms * 1e3, 2),
                              "algorithmic_GBps": round(32 * px / ms / 1e6, 1), "moved_GBps": round(48 * px / ms / 1e6, 1),
                              "frac_of_hbm_moved": round(48 * px / ms / 1e6 / hbm, 3), "hbm_peak_GBps": hbm}))
            ms = timeit(lambda: ops.warp_bwd(img, flow, go, gflow, variant), flush)
            print(json.dumps({"kernel": "warp_bwd", "variant": variant, "images": n, "us": round(ms * 1e3, 2),
                              "algorithmic_GBps": round(40 * px / ms / 1e6, 1), "moved_GBps": round(64 * px / ms / 1e6, 1),
                              "frac_of_hbm_moved": round(64 * px / ms / 1e6 / hbm, 3), "hbm_peak_GBps": hbm}))


if __name__ == "__main__":
    main()
