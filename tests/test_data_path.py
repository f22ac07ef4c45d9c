"""Data path + output writers (dsin_b200/DataProvider.py, utils.py) -- host logic, no GPU."""
import os
import types

import numpy as np
import pytest
from PIL import Image

from dsin_b200 import utils
from dsin_b200.DataProvider import Dataset, center_crop_pair, decode_png
from oracle import ms_ssim_oracle as M


def _cfg(root, crop=(32, 48), batch=2, ae_only=False, iterations=1):
    return types.SimpleNamespace(crop_size=crop, batch_size=batch, AE_only=ae_only, iterations=iterations,
                                 root_data=root, num_crops_per_img=1, do_flips=False,
                                 file_path_train="train.txt", file_path_val="val.txt", file_path_test="test.txt")


def _make_set(tmp_path, n, shape=(37, 55)):
    rng = np.random.default_rng(0)
    root, lists = str(tmp_path / "data") + os.sep, str(tmp_path / "data_paths") + os.sep
    os.makedirs(root + "image_2"), os.makedirs(root + "image_3"), os.makedirs(lists)
    imgs, lines = [], []
    for i in range(n):
        pair = []
        for cam in ("image_2", "image_3"):
            a = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
            rel = "%s/%06d_10.png" % (cam, i)
            Image.fromarray(a, "RGB").save(root + rel)
            lines.append(rel)
            pair.append(a)
        imgs.append(pair)
    for name in ("val.txt", "test.txt"):
        with open(lists + name, "w") as f:
            f.write("\n".join(lines) + "\n")
    return root, lists, imgs


def test_center_crop_offsets_and_errors():
    a = np.arange(7 * 9 * 3, dtype=np.uint8).reshape(7, 9, 3)
    x, y = center_crop_pair(a, a + 1, 4, 6)
    assert np.array_equal(x, a[1:5, 1:7]) and np.array_equal(y, (a + 1)[1:5, 1:7])  # (7-4)//2 = 1, (9-6)//2 = 1
    x, _ = center_crop_pair(a, a, 7, 9)
    assert np.array_equal(x, a)
    with pytest.raises(ValueError):
        center_crop_pair(a, a, 8, 9)
    with pytest.raises(ValueError):
        center_crop_pair(a, a[:, :8], 4, 6)


def test_decode_png_three_channels(tmp_path):
    g = np.arange(12, dtype=np.uint8).reshape(3, 4)
    Image.fromarray(g, "L").save(tmp_path / "g.png")
    assert np.array_equal(decode_png(tmp_path / "g.png"), np.repeat(g[:, :, None], 3, axis=2))
    rgba = np.dstack([np.full((3, 4), v, np.uint8) for v in (10, 20, 30, 128)])
    Image.fromarray(rgba, "RGBA").save(tmp_path / "a.png")
    assert np.array_equal(decode_png(tmp_path / "a.png"), rgba[:, :, :3])


def test_dataset_pairs_batches_and_repeat(tmp_path):
    root, lists, imgs = _make_set(tmp_path, 5)
    d = Dataset(_cfg(root, iterations=2), lists)
    val_names, test_names = d.get_data_size()
    assert len(val_names) == len(test_names) == 5
    assert test_names[3] == (root + "image_2/000003_10.png", root + "image_3/000003_10.png")
    # test batches are single pairs when the SI path is on (src/DataProvider.py:10); NCHW uint8
    for rep in range(2):
        for i in range(5):
            x, y = d.get_data_for_test()
            assert x.dtype == np.uint8 and x.shape == (1, 3, 32, 48) and y.shape == (1, 3, 32, 48)
            assert np.array_equal(x[0].transpose(1, 2, 0), imgs[i][0][2:34, 3:51])  # (37-32)//2=2, (55-48)//2=3
            assert np.array_equal(y[0].transpose(1, 2, 0), imgs[i][1][2:34, 3:51])
    with pytest.raises(StopIteration):
        d.get_data_for_test()
    # validation batches: batch_size 2, remainder dropped -> 2 batches per pass
    for b in range(4):
        x, y = d.get_data_for_val()
        assert x.shape == (2, 3, 32, 48)
        assert np.array_equal(x[1].transpose(1, 2, 0), imgs[(2 * b + 1) % 4][0][2:34, 3:51])
    with pytest.raises(NotImplementedError):
        d.get_data_for_train()
    assert Dataset(_cfg(root, ae_only=True), lists).batch_size_test == 2
    # absent lists are tolerated until iterated
    empty = Dataset(_cfg(root), str(tmp_path / "nowhere") + os.sep)
    assert empty.get_data_size() == ([], [])


def test_writers_png_name_truncation_and_loss_lists(tmp_path):
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, size=(1, 3, 176, 192)).astype(np.uint8)
    rec = np.clip(x.astype(np.float32) + rng.normal(0, 3, x.shape), 0, 255).astype(np.float32)
    y_syn = np.clip(x.astype(np.float32) + rng.normal(0, 9, x.shape), 0, 255).astype(np.float32)
    out = str(tmp_path / "images") + os.sep
    utils.save_test_imgs_fn(out, "m", rec[0], 7, 0.0312345)
    path = out + "m/7_0.03123bpp.png"
    assert os.path.isfile(path)
    assert np.array_equal(np.asarray(Image.open(path)), rec[0].transpose(1, 2, 0).astype("uint8"))  # truncation
    msssim = lambda a, b: np.float32(M.msssim_reference_call(a, b))  # noqa: E731  (CPU oracle here; CUDA kernel by default)
    os.makedirs(out, exist_ok=True)
    for _ in range(2):
        utils.loss_list_saver(x, x, rec, y_syn, 1, "m", 0.03, out, msssim_fn=msssim)
    vals = {}
    for name in ("bpp_list_", "l1_list_", "psnr_list_", "msssim_list_", "mse_list_x_y_syn_", "avg_Pearson_list_x_y_syn_"):
        lines = open(out + name + "m.txt").read().split()
        assert len(lines) == 2 and lines[0] == lines[1]
        vals[name] = float(lines[0])
    xh, rh, sh = x[0].transpose(1, 2, 0), rec[0].transpose(1, 2, 0), y_syn[0].transpose(1, 2, 0)
    assert vals["bpp_list_"] == 0.03
    assert abs(vals["l1_list_"] - np.mean(np.abs(xh.astype(np.float32) - rh))) < 1e-6
    mse = np.mean((xh.astype(np.float64) - rh.astype("uint8").astype(np.float64)) ** 2)
    assert abs(vals["psnr_list_"] - 10 * np.log10(255.0 ** 2 / mse)) < 1e-4
    assert 0.9 < vals["msssim_list_"] <= 1.0
    assert abs(vals["mse_list_x_y_syn_"] - np.mean((xh.astype(np.float32) - sh) ** 2)) < 1e-3
    # average patch Pearson against a direct evaluation
    tot, n = 0.0, 0
    for i in range(0, 176 - 19, 20):
        for j in range(0, 192 - 23, 24):
            a, b = xh[i:i + 20, j:j + 24].reshape(-1).astype(np.float64), sh[i:i + 20, j:j + 24].reshape(-1).astype(np.float64)
            tot += np.corrcoef(a, b)[0, 1]
            n += 1
    assert abs(vals["avg_Pearson_list_x_y_syn_"] - tot / n) < 1e-6
