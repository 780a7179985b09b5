"""CPU restatement of the reference's ``Super`` loss (test infrastructure; reference loss.py:246-274).

    loss = 204 * L1(sr, hr)
         + 102 * [L1(g_I0_F_t_0, hr) + L1(g_I1_F_t_1, hr) + L1(warp(I0, F_1_0), I1) + L1(warp(I1, F_0_1), I0)]
         + 0.005 * MSE(vgg16_conv4_3(sr), vgg16_conv4_3(hr))
         + smooth(F_1_0) + smooth(F_0_1),   smooth(F) = mean|F[..., :-1] - F[..., 1:]| + mean|F[..., :-1, :] - F[..., 1:, :]|

``vgg16_conv4_3`` is torchvision's ``vgg16().features[:22]`` (loss.py:249-250): ten 3x3 convolutions with ReLU after
all but the last, max-pools after the 2nd, 4th and 7th.  The reference downloads the ImageNet weights; they are not
available offline, so parity is pinned with a seeded random-initialised VGG16 handed to both sides
(tests/test_oracle.py::test_super_loss_equals_live_reference).
"""
import torch
import torch.nn.functional as F

# (index in torchvision's vgg16.features, cin, cout); 'M' = 2x2 max-pool
VGG16_CONV4_3 = [(0, 3, 64), (2, 64, 64), 'M', (5, 64, 128), (7, 128, 128), 'M', (10, 128, 256), (12, 256, 256),
                 (14, 256, 256), 'M', (17, 256, 512), (19, 512, 512), (21, 512, 512)]


def seeded_vgg16_state(seed=0):
    """Random conv4_3 weights in torchvision's key names (``features.N.weight/bias``), He-scaled so features keep O(1)
    magnitude through the ten layers."""
    g = torch.Generator().manual_seed(seed)
    state = {}
    for e in VGG16_CONV4_3:
        if e == 'M':
            continue
        idx, cin, cout = e
        state["features.%d.weight" % idx] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
        state["features.%d.bias" % idx] = torch.randn(cout, generator=g) * 0.05
    return state


def vgg16_conv4_3(x, state):
    last = [e for e in VGG16_CONV4_3 if e != 'M'][-1][0]
    for e in VGG16_CONV4_3:
        if e == 'M':
            x = F.max_pool2d(x, 2, 2)
            continue
        idx = e[0]
        x = F.conv2d(x, state["features.%d.weight" % idx], state["features.%d.bias" % idx], padding=1)
        if idx != last:
            x = F.relu(x)
    return x


def smoothness(flow):
    return (flow[:, :, :, :-1] - flow[:, :, :, 1:]).abs().mean() + (flow[:, :, :-1, :] - flow[:, :, 1:, :]).abs().mean()


def super_loss(sr, hr, aux, i0, i1, vgg_state):
    """``aux`` = the dict the SuperSloMo plugin returns next to the prediction (superslomo/model.py:640-643)."""
    f01, f10 = aux["bidirectional_flow"]
    g0, g1 = aux["warped_intermediate_frames"]
    w0, w1 = aux["warped_input_frames"]
    l1 = lambda a, b: (a - b).abs().mean()
    recn = l1(sr, hr)
    prcp = ((vgg16_conv4_3(sr, vgg_state) - vgg16_conv4_3(hr, vgg_state)) ** 2).mean()
    warp = l1(g0, hr) + l1(g1, hr) + l1(w0, i1) + l1(w1, i0)
    smooth = smoothness(f10) + smoothness(f01)
    return 204 * recn + 102 * warp + 0.005 * prcp + smooth
