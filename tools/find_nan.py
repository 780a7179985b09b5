"""Debugging aid: run one forward/backward of a backbone with NaN-poisoned fresh buffers (MI_B200_POISON=1) and
report the first operator call whose logical output contains NaN (= a kernel that read what nobody wrote)."""
import os
import sys

os.environ["MI_B200_POISON"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_args  # noqa: E402
from meta_interpolation_b200 import backbone  # noqa: E402
from meta_interpolation_b200.meta_learning_system import _build_backbone  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "rrin"
hw = (64, 128) if len(sys.argv) < 4 else (int(sys.argv[2]), int(sys.argv[3]))
ops = backbone.default_ops()
reported = [0]


def wrap(name, fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        cands = [("ret", r)] + [("arg%d" % i, t) for i, t in enumerate(a)] + list(k.items())
        for tag, t in cands:
            if torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.numel() > 0:
                if torch.isnan(t).any().item() and reported[0] < 12:
                    reported[0] += 1
                    print("NaN after %-22s in %-6s shape %s strides %s" % (name, tag, tuple(t.shape), t.stride()))
        return r
    return inner


for n in dir(ops):
    if n.startswith("_") or n in ("empty_act", "zeros_act", "empty_like_act", "empty_weight", "workspace",
                                  "launch_count", "set_workspace_slot", "prof_enable", "prof_summary"):
        continue
    f = getattr(ops, n)
    if callable(f):
        setattr(ops, n, wrap(n, f))

net = _build_backbone(make_args(model=model, cuda=True), ops)
g = torch.Generator().manual_seed(0)
f0, f1, t = (torch.rand(1, 3, *hw, generator=g).cuda() for _ in range(3))
fast = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
out = net.forward(f0, f1, params=fast)
if isinstance(out, tuple):
    out = out[0]
print("forward NaN:", torch.isnan(out).any().item())
gr = torch.autograd.grad((out - t).abs().mean(), list(fast.values()), allow_unused=True)
print("grad NaN:", [k for k, g_ in zip(fast, gr) if g_ is not None and torch.isnan(g_).any().item()][:8])
