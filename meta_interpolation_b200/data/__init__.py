"""Septuplet provider behind the interface ``ExperimentBuilder`` drives (reference data/__init__.py:520-640,
experiment_builder.py:228-263): ``MetaLearningSystemDataLoader(args).get_train_batches() / get_val_batches() /
get_test_batches()`` yield ``(frames, metadata)`` with ``frames`` = the 7 float tensors [B,3,h,w] of a meta-batch,
already resident on the GPU.

DataLoader workers only decode; crop, temporal flip, BGR->RGB, /255 and the per-model normalisation of the whole
meta-batch are one kernel (``vimeo_septuplet.py`` / ``csrc/data.cu``) after a uint8 host-to-device copy.  Only the
Vimeo-90K septuplets are on the hot path; the reference's evaluation-only sets (Middlebury, HD, DAVIS, SNU-FILM, raw
video) are outside SURVEY section 8.
"""
from torch.utils.data import DataLoader

from .vimeo_septuplet import VimeoSeptuplet

_SPLITS = ("train", "val", "test")


class MetaLearningSystemDataLoader(object):
    def __init__(self, args, current_iter=0, ops=None):
        if args.dataset != 'vimeo90k':
            raise NotImplementedError('dataset %s is outside the B200 hot path (SURVEY section 8f)' % args.dataset)
        self.args = args
        self.num_of_gpus = args.num_gpu
        self.num_workers = args.num_workers
        self.batch_size = dict(zip(_SPLITS, (args.batch_size, args.val_batch_size, args.test_batch_size)))
        self.dataset = VimeoSeptuplet(args=args, ops=ops)
        self.full_data_length = self.dataset.data_length
        # how many training samples earlier epochs / a resumed run already consumed (reference :568-574): it seeds
        # the training split's shuffling and augmentation in ``switch_set``
        self.total_train_iters_produced = 0
        self.continue_from_iter(current_iter)

    def continue_from_iter(self, current_iter):
        self.total_train_iters_produced += current_iter * self.batch_size["train"]

    def get_dataloader(self, mode='train'):
        """Stock DataLoader over the split (shuffled for training only, short last batch kept: reference :559-566);
        pinned so the uint8 frames go to the device with one asynchronous copy."""
        on_gpu = self.dataset.ops is None or self.dataset.ops.name == 'cuda'
        return DataLoader(self.dataset, batch_size=self.batch_size[mode], shuffle=mode == "train", drop_last=False,
                          num_workers=self.num_workers, pin_memory=on_gpu)

    def _stream(self, split, total_batches):
        """One pass over ``split``: optionally cap its length at ``total_batches`` meta-batches, point the dataset
        at the split, then stage every decoded batch on the device."""
        ds = self.dataset
        if total_batches == -1:
            ds.data_length = self.full_data_length
        else:
            ds.data_length[split] = total_batches * ds.batch_size[split]
        if split == "train":
            ds.switch_set(set_name=split, current_iter=self.total_train_iters_produced)
            self.total_train_iters_produced += self.batch_size[split]
        else:
            ds.switch_set(set_name=split)
        for decoded, metadata in self.get_dataloader(mode=split):
            yield ds.to_device(decoded), metadata

    def get_train_batches(self, total_batches=-1, augment_images=False):
        return self._stream("train", total_batches)

    def get_val_batches(self, total_batches=-1, augment_images=False):
        return self._stream("val", total_batches)

    def get_test_batches(self, total_batches=-1, augment_images=False):
        return self._stream("test", total_batches)
