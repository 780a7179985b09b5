"""Input/output padding geometry shared by the plugins (reference model_utils.py:17-28,
superslomo/model.py:567-575, voxelflow/core/models/voxel_flow.py:360-368)."""


def reflect_pads(height, width, shift):
    """(left, right, top, bottom) reflection pads up to a multiple of ``2**shift``; floor on the left/top."""
    pw = ph = 0
    if width != ((width >> shift) << shift):
        pw = (((width >> shift) + 1) << shift) - width
    if height != ((height >> shift) << shift):
        ph = (((height >> shift) + 1) << shift) - height
    return pw // 2, pw - pw // 2, ph // 2, ph - ph // 2


def xavier_or_zero(name, shape):
    """Default MetaConv2dLayer init (reference model_utils.py:329-333)."""
    import torch
    if name.endswith(".weight") and len(shape) == 4:
        w = torch.empty(*shape)
        torch.nn.init.xavier_uniform_(w)
        return w
    return torch.zeros(*shape)
