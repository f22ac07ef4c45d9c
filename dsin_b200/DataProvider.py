"""Validation / test side of the reference's data path (/root/reference/src/DataProvider.py) without
tf.data.

    Dataset(configs, current_directory)          src/DataProvider.py:5-20
    get_data_size() -> (val pairs, test pairs)   src/DataProvider.py:186-187
    get_data_for_val() / get_data_for_test()     src/DataProvider.py:193-199  -> [x, y] uint8 (B,3,H,W)

A pair list holds alternating lines "x path", "y path", each prefixed with `configs.root_data`
(src/DataProvider.py:96-100,166-170).  Both images are decoded to 3 channels, concatenated, centre-cropped to
`crop_size` with offsets (size - crop) // 2 (src/DataProvider.py:62-94) and batched with drop_remainder; the
test batch is 1 unless AE_only (src/DataProvider.py:10).  The list is walked `configs.iterations` times like the
reference's `.repeat(count=iterations)`.  The training pipeline (random crops, flips, shuffling) belongs to the
training path (N4) and raises NotImplementedError.  Host-side only; decoding uses PIL.
"""
from __future__ import annotations

import os

import numpy as np


def decode_png(path):
    """tf.image.decode_png(..., channels=3) (src/DataProvider.py:26-27): uint8 HWC, grey replicated, alpha dropped."""
    from PIL import Image
    with Image.open(path) as im:
        if im.mode in ("I;16", "I;16B", "I"):  # 16-bit PNG decoded as uint8 keeps the high byte
            a = np.asarray(im).astype(np.uint32)
            a = (a >> 8).astype(np.uint8)
            return np.repeat(a[:, :, None], 3, axis=2)
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def center_crop_pair(x_img, y_img, crop_h, crop_w):
    """si_opt_crop_img with crop == crop_size (src/DataProvider.py:62-94): one centre crop of x and y."""
    if x_img.shape != y_img.shape:
        raise ValueError("x and y images differ in shape: %s vs %s" % (x_img.shape, y_img.shape))
    H, W = x_img.shape[:2]
    if H < crop_h or W < crop_w:
        raise ValueError("image %dx%d is smaller than the crop %dx%d" % (H, W, crop_h, crop_w))
    oh, ow = (H - crop_h) // 2, (W - crop_w) // 2
    return x_img[oh:oh + crop_h, ow:ow + crop_w], y_img[oh:oh + crop_h, ow:ow + crop_w]


class Dataset(object):
    def __init__(self, configs, current_directory, buffer_size_param=50, num_parallel_calls=6):
        self.crop_size_h, self.crop_size_w = configs.crop_size[0], configs.crop_size[1]
        self.batch_size = configs.batch_size
        self.batch_size_test = configs.batch_size if configs.AE_only else 1
        self.iterations = configs.iterations
        self.root_data = configs.root_data
        self.file_path_train = current_directory + configs.file_path_train
        self.file_path_val = current_directory + configs.file_path_val
        self.file_path_test = current_directory + configs.file_path_test
        self.loadData()

    def readfiles(self, fname):
        with open(fname) as f:
            return [self.root_data + line.strip() for line in f.readlines()]

    @staticmethod
    def _pairs(content):
        return list(zip(content[0:][::2], content[1:][::2]))

    def loadData(self):
        # the inference drop-in tolerates absent lists (only the splits that are iterated must exist)
        self.val_imgs_names = self._pairs(self.readfiles(self.file_path_val)) \
            if os.path.isfile(self.file_path_val) else []
        self.test_imgs_names = self._pairs(self.readfiles(self.file_path_test)) \
            if os.path.isfile(self.file_path_test) else []
        self._val_iter = self._batches(self.val_imgs_names, self.batch_size)
        self._test_iter = self._batches(self.test_imgs_names, self.batch_size_test)

    def _load_pair(self, names):
        x, y = center_crop_pair(decode_png(names[0]), decode_png(names[1]), self.crop_size_h, self.crop_size_w)
        return x, y

    def _batches(self, pairs, batch):
        for _ in range(self.iterations):
            for b in range(len(pairs) // batch):  # drop_remainder=True
                xs, ys = zip(*(self._load_pair(p) for p in pairs[b * batch:(b + 1) * batch]))
                yield [np.ascontiguousarray(np.transpose(np.stack(xs), (0, 3, 1, 2))),
                       np.ascontiguousarray(np.transpose(np.stack(ys), (0, 3, 1, 2)))]

    def get_data_size(self):
        return self.val_imgs_names, self.test_imgs_names

    def get_data_for_train(self):
        raise NotImplementedError("dsin_b200 implements the inference path only (training data pipeline: SURVEY 8f N4)")

    def get_data_for_val(self):
        return next(self._val_iter)

    def get_data_for_test(self):
        return next(self._test_iter)
