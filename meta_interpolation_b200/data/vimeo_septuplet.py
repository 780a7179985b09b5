"""Vimeo-90K septuplets staged for the inner loop (drop-in for the reference's ``data/vimeo_septuplet.py``).

Same constructor arguments, list files, attributes (``crop_size``, ``frames``, ``current_set_name``,
``data_length``), ``switch_set`` / ``__len__`` and -- under the same ``random`` seed -- the same crop origins and
temporal flips as reference data/vimeo_septuplet.py:10-92, because the three ``random`` draws are made in the
reference's order (:57-58, :64).

What differs is WHERE the pixel work happens.  The reference crops, flips, re-orders channels, converts to float and
normalises inside the DataLoader workers and ships 7 float frames per task over PCIe.  Here ``__getitem__`` only
decodes the PNGs and records the augmentation decisions; the uint8 frames are collated into one pinned buffer, cross
PCIe once (4x fewer bytes) and ``to_device`` turns the whole meta-batch into the 7 float NCHW tensors
``run_train_iter`` expects with ONE kernel launch (``mi_septuplet_prepare``).
"""
import os
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from ..backbone import default_ops


class VimeoSeptuplet(Dataset):
    NORMALISATION = {      # reference :31-40 (transforms.Normalize(mean, std) applied after the float conversion)
        'superslomo': ([0.429, 0.431, 0.397], [1.0, 1.0, 1.0]),
        'voxelflow': ([0.5 * 255] * 3, [0.5 * 255] * 3),
    }

    def __init__(self, args, ops=None):
        self.args = args
        self.ops = ops
        self.data_root = args.data_root
        self.image_root = os.path.join(self.data_root, 'sequences')
        with open(os.path.join(self.data_root, 'sep_trainlist.txt'), 'r') as f:
            self.trainlist = f.read().splitlines()
        with open(os.path.join(self.data_root, 'sep_testlist.txt'), 'r') as f:
            self.testlist = f.read().splitlines()
        self.batch_size = {'train': args.batch_size, 'val': args.val_batch_size, 'test': args.test_batch_size}
        self.crop_size = 256
        self.frames = [1, 2, 3, 4, 5, 6, 7]
        self.current_set_name = "train" if args.mode == 'train' else 'val'
        self.data_length = {'train': len(self.trainlist), 'val': len(self.testlist), 'test': 0}
        self.mean, self.std = self.NORMALISATION.get(args.model, (None, None))
        self.div255 = args.model != 'voxelflow'      # reference :72-75

    @staticmethod
    def _decode(path):
        import cv2                   # the reference's decoder (BGR, :52); the kernel swaps the channels
        image = cv2.imread(path)
        if image is None:
            raise FileNotFoundError(path)
        return image

    def __getitem__(self, index):
        names = self.trainlist if self.current_set_name == 'train' else self.testlist
        imgpath = os.path.join(self.image_root, names[index % len(names)])
        imgpaths = ['%s/im%d.png' % (imgpath, i) for i in self.frames]
        raw = np.stack([self._decode(p) for p in imgpaths])          # [7, H, W, 3] uint8, BGR
        H, W = raw.shape[1:3]
        y0, x0, h, w, flip = 0, 0, H, W, False
        if self.current_set_name == 'train':
            y0 = random.randint(0, max(0, H - self.crop_size))
            x0 = random.randint(0, max(0, W - self.crop_size))
            h, w = min(self.crop_size, H - y0), min(self.crop_size, W - x0)   # numpy slicing clips (:59-61)
            if random.random() >= 0.5:
                flip = True
                imgpaths = imgpaths[::-1]
        staged = {'raw': torch.from_numpy(raw), 'y0': y0, 'x0': x0, 'h': h, 'w': w, 'reversed': flip}
        return staged, {'imgpaths': imgpaths}

    def to_device(self, staged, device=None):
        """Collated ``staged`` dict (``raw`` [B,7,H,W,3] uint8, per-task ``y0/x0/h/w/reversed``) -> the list of 7 float
        tensors [B,3,h,w] the meta system consumes.  One host-to-device copy of the uint8 batch + one launch."""
        ops = self.ops if self.ops is not None else default_ops()
        device = torch.device(device) if device is not None else ops.device
        raw = staged['raw']
        tasks, _, H, W, _ = raw.shape
        h, w = int(staged['h'][0]), int(staged['w'][0])
        y0 = torch.as_tensor(staged['y0'], dtype=torch.int32).reshape(-1)
        x0 = torch.as_tensor(staged['x0'], dtype=torch.int32).reshape(-1)
        rev = torch.as_tensor(staged['reversed']).reshape(-1).to(torch.uint8)
        if not (bool((torch.as_tensor(staged['h']) == h).all()) and bool((torch.as_tensor(staged['w']) == w).all())):
            raise ValueError('the tasks of one meta-batch must share one crop size')
        if int(y0.min()) < 0 or int(x0.min()) < 0 or int(y0.max()) + h > H or int(x0.max()) + w > W:
            raise ValueError('crop window outside the %dx%d source frames' % (H, W))
        nb = ops.name == 'cuda'
        out = ops.septuplet_prepare(raw.contiguous().to(device, non_blocking=nb), y0.to(device, non_blocking=nb),
                                    x0.to(device, non_blocking=nb), rev.to(device, non_blocking=nb), h, w, bgr=True,
                                    div255=self.div255, mean=self.mean, std=self.std)
        return [out[f] for f in range(out.shape[0])]

    def switch_set(self, set_name, current_iter=None):
        self.current_set_name = set_name

    def __len__(self):
        return self.data_length[self.current_set_name]
