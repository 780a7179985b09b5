"""Flat parameter arenas.

One contiguous fp32 buffer holds every tensor of a backbone at a fixed,
1024-float-aligned offset (decision 1).  Conv weights are
stored KRSC with Cin padded to 4 (pad lanes zero) so the tensor-core kernels can
address them with TMA; the reference-visible ``nn.Parameter`` of shape
[Cout,Cin,k,k] is a permuted VIEW of the same storage (``named_parameters()``,
``state_dict()`` and ``load_state_dict()`` therefore keep the reference's key
names and shapes with no copy).  The same layout is reused for fast weights,
gradients, Meta-SGD alphas and optimizer moments, so the inner update, the
outer optimizer and the NCCL all-reduce each touch one flat buffer.
"""
from collections import OrderedDict

import torch

CHUNK = 1024


def pad4(c):
    return (c + 3) & ~3


class Layout:
    """name -> (offset, logical shape) for one backbone; shared by all arenas of that backbone."""

    def __init__(self, named_shapes):
        self.entries = OrderedDict()
        off = 0
        for idx, (name, shape) in enumerate(named_shapes):
            shape = tuple(int(s) for s in shape)
            if len(shape) == 4:
                cout, cin, k, _ = shape
                numel = cout * k * k * pad4(cin)
            else:
                numel = 1
                for s in shape:
                    numel *= s
            self.entries[name] = (off, shape, numel, idx)
            off += (numel + CHUNK - 1) // CHUNK * CHUNK
        self.total = off
        self.names = list(self.entries.keys())

    def index(self, name):
        return self.entries[name][3]

    def segment_table(self):
        """int32 [total/1024]: tensor index owning each 1024-float chunk (-1 = padding)."""
        seg = torch.full((self.total // CHUNK,), -1, dtype=torch.int32)
        for name, (off, shape, numel, idx) in self.entries.items():
            seg[off // CHUNK:(off + numel + CHUNK - 1) // CHUNK] = idx
        return seg

    def logical_numel(self, name):
        n = 1
        for s in self.entries[name][1]:
            n *= s
        return n


class Arena:
    """One flat buffer laid out by ``Layout``."""

    def __init__(self, layout, device, data=None):
        self.layout = layout
        self.flat = data if data is not None else torch.zeros(layout.total, device=device, dtype=torch.float32)
        self._views = {}

    def kernel_view(self, name):
        """KRSC view [Cout,k,k,Cin] (stride(2)=pad4(Cin)) for conv weights, logical view otherwise."""
        v = self._views.get(name)
        if v is None:
            off, shape, numel, _ = self.layout.entries[name]
            if len(shape) == 4:
                cout, cin, k, _ = shape
                v = self.flat[off:off + numel].view(cout, k, k, pad4(cin))[..., :cin]
            else:
                v = self.flat[off:off + numel].view(shape)
            self._views[name] = v
        return v

    def reference_view(self, name):
        """View with the reference's shape: OIHW for conv weights (permuted, non-contiguous)."""
        v = self.kernel_view(name)
        return v.permute(0, 3, 1, 2) if v.dim() == 4 else v

    def zero_(self):
        self.flat.zero_()
        return self

    def copy_from(self, other):
        self.flat.copy_(other.flat)
        return self
