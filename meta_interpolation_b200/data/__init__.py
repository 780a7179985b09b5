"""Data provider with the reference's interface (reference data/__init__.py:520-640) for the Vimeo-90K septuplets.

``MetaLearningSystemDataLoader(args).get_train_batches()/get_val_batches()/get_test_batches()`` yield
``(frames, metadata)`` exactly where ``ExperimentBuilder`` expects them (experiment_builder.py:228-263), with
``frames`` already on the GPU: a list of 7 float tensors [B,3,h,w].  DataLoader workers only decode; the pixel work is
one kernel per meta-batch (see ``vimeo_septuplet.py``).  The other datasets of the reference (Middlebury, HD, DAVIS,
SNU-FILM, raw video) are evaluation-only and outside SURVEY section 8.
"""
from torch.utils.data import DataLoader

from .vimeo_septuplet import VimeoSeptuplet


class MetaLearningSystemDataLoader(object):
    def __init__(self, args, current_iter=0, ops=None):
        self.num_of_gpus = args.num_gpu
        self.batch_size = {'train': args.batch_size, 'val': args.val_batch_size, 'test': args.test_batch_size}
        self.num_workers = args.num_workers
        self.total_train_iters_produced = 0
        if args.dataset != 'vimeo90k':
            raise NotImplementedError('dataset %s is outside the B200 hot path (SURVEY section 8f)' % args.dataset)
        self.dataset = VimeoSeptuplet(args=args, ops=ops)
        self.full_data_length = self.dataset.data_length
        self.continue_from_iter(current_iter=current_iter)
        self.args = args

    def get_dataloader(self, mode='train'):
        pin = self.dataset.ops is None or self.dataset.ops.name == 'cuda'
        return DataLoader(self.dataset, batch_size=self.batch_size[mode], shuffle=(mode == 'train'),
                          num_workers=self.num_workers, drop_last=False, pin_memory=pin)

    def continue_from_iter(self, current_iter):
        self.total_train_iters_produced += (current_iter * self.batch_size["train"])

    def _batches(self, mode):
        for staged, metadata in self.get_dataloader(mode=mode):
            yield self.dataset.to_device(staged), metadata

    def get_train_batches(self, total_batches=-1, augment_images=False):
        if total_batches == -1:
            self.dataset.data_length = self.full_data_length
        else:
            self.dataset.data_length["train"] = total_batches * self.dataset.batch_size["train"]
        self.dataset.switch_set(set_name="train", current_iter=self.total_train_iters_produced)
        self.total_train_iters_produced += self.batch_size["train"]
        yield from self._batches("train")

    def get_val_batches(self, total_batches=-1, augment_images=False):
        if total_batches == -1:
            self.dataset.data_length = self.full_data_length
        else:
            self.dataset.data_length['val'] = total_batches * self.dataset.batch_size["val"]
        self.dataset.switch_set(set_name="val")
        yield from self._batches("val")

    def get_test_batches(self, total_batches=-1, augment_images=False):
        if total_batches == -1:
            self.dataset.data_length = self.full_data_length
        else:
            self.dataset.data_length['test'] = total_batches * self.dataset.batch_size["test"]
        self.dataset.switch_set(set_name='test')
        yield from self._batches("test")
