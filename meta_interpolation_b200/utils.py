"""Metrics helpers with the reference's names and arithmetic (reference utils.py:135-204).

PSNR is computed on 8-bit-quantised tensors with ``mse + 1e-8`` exactly as
utils.py:171-186,195-204; the squared-error reduction runs in one sm_100a kernel
(``mi_psnr_accumulate``).  SSIM follows pytorch_msssim/__init__.py:19-75 (11x11
gaussian, valid convolution) and is a logging-only metric.

Checkpoint helpers (reference utils.py:34-118, 253-255) keep the reference's file layout
(``checkpoint/<exp_name>/checkpoint.pth`` + ``model_best.pth``; a dict with ``epoch``, ``arch``, ``state_dict``,
``best_PSNR`` and optionally ``optimizer``) and its lossy matching rules, so files written by either code base load in
the other: the state-dict keys and shapes of this package's modules are the reference's.
"""
import math
import os
import shutil

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


_WINDOWS = {}


def gaussian_1d(size):
    """pytorch_msssim/__init__.py:7-9: normalised float32 Gaussian (sigma 1.5) of ``size`` taps, on the host."""
    g = _WINDOWS.get(size)
    if g is None:
        g = torch.Tensor([math.exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])
        g = (g / g.sum()).contiguous()
        _WINDOWS[size] = g
    return g


def calc_metrics_device(ops, im_pred, im_gt):
    """PSNR and SSIM of one [3,H,W] pair with two kernels and ONE host read (utils.py:195-204)."""
    pred, gt = im_pred.detach().contiguous(), im_gt.detach().contiguous()
    c, h, w = pred.shape
    win = min(11, h, w)
    acc = torch.zeros(2, dtype=torch.float64, device=pred.device)
    ops.psnr_accumulate(pred, gt, acc[0:1])
    ops.ssim_accumulate(pred, gt, gaussian_1d(win), acc[1:2], 255.0)
    sq, ss = acc.tolist()
    psnr = -10 * math.log10(sq / pred.numel() + 1e-8)
    return psnr, torch.tensor(ss / (c * (h - win + 1) * (w - win + 1)), dtype=torch.float32, device=pred.device)


def calc_metrics(im_pred, im_gt, ops=None):
    """reference utils.py:195-204 -> (psnr, ssim) for one [3,H,W] pair in [0,1]."""
    if ops is not None and ops.name == "cuda":
        return calc_metrics_device(ops, im_pred, im_gt)
    d = (quantize(im_pred.detach(), 1.) - quantize(im_gt.detach(), 1.)).div(255)
    psnr = -10 * math.log10(float(d.pow(2).mean()) + 1e-8)
    s = ssim(quantize(im_pred.detach(), 1.).unsqueeze(0), quantize(im_gt.detach(), 1.).unsqueeze(0), val_range=255)
    return psnr, s


# ---------------------------------------------------------------------- checkpoints (reference utils.py:34-118)
def _matching_entries(own_state, ckpt_state, report_missing_keys):
    """Entries of ``ckpt_state`` whose key exists in ``own_state`` with the same shape; second value is True when
    anything was left out (a shape differs, or -- load_checkpoint only, :47-58 -- a key is unknown / absent)."""
    picked, partial = {}, False
    for key, value in ckpt_state.items():
        if key not in own_state:
            partial = partial or report_missing_keys
            continue
        if own_state[key].size() != value.size():
            print('Size mismatch while loading!   %s != %s   Skipping %s...'
                  % (str(own_state[key].size()), str(value.size()), key))
            partial = True
            continue
        picked[key] = value
    if report_missing_keys and len(own_state) > len(picked):
        partial = True
    return picked, partial


def _read(path, device=None):
    # the reference's files pickle the argparse Namespace under 'arch' (experiment_builder.py:308-314)
    return torch.load(path, map_location=device, weights_only=False)


def update_lr(optimizer, lr):
    """reference utils.py:253-255."""
    for group in optimizer.param_groups:
        group['lr'] = lr


def load_checkpoint(args, model, optimizer, fix_loaded=False):
    """reference utils.py:34-86: resume ``model`` (and, when nothing was skipped, ``optimizer``) from
    ``checkpoint/<resume_exp>/checkpoint.pth`` (``model_best.pth`` in val / test mode); sets ``args.start_epoch``."""
    if args.resume_exp is None:
        args.resume_exp = args.exp_name
    fname = 'model_best.pth' if args.mode in ['val', 'test'] else 'checkpoint.pth'
    load_name = os.path.join('checkpoint', args.resume_exp, fname)
    print("loading checkpoint %s" % load_name)
    checkpoint = _read(load_name, getattr(model, 'device', None))
    args.start_epoch = checkpoint['epoch'] if args.resume_exp == args.exp_name else 0
    own = model.state_dict()
    picked, partial = _matching_entries(own, checkpoint['state_dict'], report_missing_keys=True)
    own.update(picked)
    model.load_state_dict(own)
    if not partial and optimizer is not None and args.resume_exp is not None and args.mode != 'test':
        optimizer.load_state_dict(checkpoint['optimizer'])
        update_lr(optimizer, args.lr)
    if fix_loaded:
        for key, param in model.named_parameters():
            if key in picked:
                print(key)
                param.requires_grad = False
    print("loaded checkpoint %s" % load_name)


def lossy_load_state_dict(net, ckpt_state_dict, opt=None, ckpt_optimizer=None):
    """reference utils.py:89-107: copy every entry that fits, ignore the rest."""
    own = net.state_dict()
    picked, partial = _matching_entries(own, ckpt_state_dict, report_missing_keys=False)
    own.update(picked)
    net.load_state_dict(own)
    if opt is not None and not partial:
        opt.load_state_dict(ckpt_optimizer)


def save_checkpoint(state, is_best, exp_name, filename='checkpoint.pth'):
    """reference utils.py:110-118."""
    directory = os.path.join('checkpoint', exp_name)
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, filename)
    torch.save(state, path)
    if is_best:
        shutil.copyfile(path, os.path.join(directory, 'model_best.pth'))
