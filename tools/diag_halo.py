"""GPU diagnostic: per-tap error of the tcgen05 conv engines against the fp32 SIMT engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402
from meta_interpolation_b200.ops import ENGINE_SIMT, ENGINE_TC  # noqa: E402


def main():
    ops = default_ops()
    print("HALO=%s HALO_BO=%s" % (os.environ.get("MI_B200_HALO"), os.environ.get("MI_B200_HALO_BO")))
    for (n, h, w, cin, cout) in [(1, 16, 8, 32, 32), (1, 16, 8, 64, 64), (1, 40, 24, 32, 32)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = ops.empty_act(n, h, w, cin)
        x.copy_(torch.rand(n, h, w, cin, device="cuda", generator=g) - 0.5)
        errs = []
        for tap in range(9):
            wt = ops.empty_weight(cout, cin, 3)
            wt[:, tap // 3, tap % 3, :] = torch.rand(cout, cin, device="cuda", generator=g) - 0.5
            ys = ops.conv_fprop(x, wt, None, engine=ENGINE_SIMT)
            yt = ops.conv_fprop(x, wt, None, engine=ENGINE_TC)
            torch.cuda.synchronize()
            errs.append(float((ys - yt).abs().max() / ys.abs().max()))
        print("shape", (n, h, w, cin, cout), "rel err per tap:", " ".join("%.1e" % e for e in errs))


if __name__ == "__main__":
    main()
