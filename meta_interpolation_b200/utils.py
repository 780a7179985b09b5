"""Metrics helpers with the reference's names and arithmetic (reference utils.py:135-204).

PSNR is computed on 8-bit-quantised tensors with ``mse + 1e-8`` exactly as
utils.py:171-186,195-204; the squared-error reduction runs in one sm_100a kernel
(``mi_psnr_accumulate``).  SSIM follows pytorch_msssim/__init__.py:19-75 (11x11
gaussian, valid convolution) and is a logging-only metric.
"""
import math

import torch
import torch.nn.functional as F


class AverageMeter(object):
    """Computes and stores the average and current value (reference utils.py:135-150)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def quantize(img, rgb_range=255):
    return img.mul(255 / rgb_range).clamp(0, 255).round()


def calc_psnr_device(ops, pred, gt):
    """PSNR of two [0,1] images via the fused quantise+MSE kernel; one host read."""
    sq = torch.zeros(1, dtype=torch.float64, device=pred.device)
    ops.psnr_accumulate(pred.contiguous(), gt.contiguous(), sq)
    mse = float(sq.item()) / pred.numel() + 1e-8
    return -10 * math.log10(mse)


def _gaussian_window(size, channel, device):
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, size, size).contiguous().to(device)


def ssim(img1, img2, window_size=11, val_range=255):
    """pytorch_msssim/__init__.py:19-75 with size_average=True."""
    _, channel, height, width = img1.size()
    real = min(window_size, height, width)
    window = _gaussian_window(real, channel, img1.device)
    mu1 = F.conv2d(img1, window, groups=channel)
    mu2 = F.conv2d(img2, window, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, window, groups=channel) - mu1_sq
    s2 = F.conv2d(img2 * img2, window, groups=channel) - mu2_sq
    s12 = F.conv2d(img1 * img2, window, groups=channel) - mu1_mu2
    c1, c2 = (0.01 * val_range) ** 2, (0.03 * val_range) ** 2
    v1, v2 = 2.0 * s12 + c2, s1 + s2 + c2
    return (((2 * mu1_mu2 + c1) * v1) / ((mu1_sq + mu2_sq + c1) * v2)).mean()


def calc_metrics(im_pred, im_gt, ops=None):
    """reference utils.py:195-204 -> (psnr, ssim) for one [3,H,W] pair in [0,1]."""
    if ops is not None and ops.name == "cuda":
        psnr = calc_psnr_device(ops, im_pred.detach(), im_gt.detach())
    else:
        d = (quantize(im_pred.detach(), 1.) - quantize(im_gt.detach(), 1.)).div(255)
        psnr = -10 * math.log10(float(d.pow(2).mean()) + 1e-8)
    s = ssim(quantize(im_pred.detach(), 1.).unsqueeze(0), quantize(im_gt.detach(), 1.).unsqueeze(0), val_range=255)
    return psnr, s
