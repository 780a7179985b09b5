"""Run a few launches of one conv shape through the tcgen05 engine (target for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meta_interpolation_b200.backbone import default_ops  # noqa: E402
from meta_interpolation_b200.ops import ENGINE_TC, WG_STORE, WgradSpec, pad4  # noqa: E402

n, h, w, cin, cout = [int(v) for v in sys.argv[1:6]]
mode = sys.argv[6] if len(sys.argv) > 6 else "fprop"
ops = default_ops()
x = ops.empty_act(n, h, w, cin); x.copy_(torch.rand(n, h, w, cin, device="cuda") - 0.5)
dy = ops.empty_act(n, h, w, cout); dy.copy_(torch.rand(n, h, w, cout, device="cuda") - 0.5)
wt = ops.empty_weight(cout, cin, 3); wt.copy_(torch.rand(cout, 3, 3, cin, device="cuda") - 0.5)
b = torch.rand(cout, device="cuda")
y = ops.empty_act(n, h, w, cout)
gw, gb = ops.empty_weight(cout, cin, 3), torch.zeros(cout, device="cuda")
for _ in range(4):
    if mode == "fprop":
        ops.conv_fprop(x, wt, b, 1, 0.0, out=y, engine=ENGINE_TC)
    else:
        ops.conv_wgrad(x, dy, 3, pad4(cin), WgradSpec(WG_STORE, grad_w=gw, grad_b=gb), engine=ENGINE_TC)
torch.cuda.synchronize()
