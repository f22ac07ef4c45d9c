"""Metric helpers and output writers of the reference's utils.py (/root/reference/src/utils.py:82-180) on
libdsin_b200.  Plotting (utils.py:12-79) is presentation code and out of scope."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops


def l1_x_vs_rec(x, x_rec):
    """L1 between original and reconstruction (utils.py:82-87)."""
    diff = np.absolute(x.astype("float32") - x_rec.astype("float32"))
    return diff.astype("uint8"), np.mean(diff)


def psnr_x_vs_rec(x, x_rec):
    """skimage.measure.compare_psnr(x, uint8(x_rec)) for uint8 data range (utils.py:90-91)."""
    a = np.asarray(x).astype(np.float64)
    b = np.asarray(x_rec).astype("uint8").astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return np.float32(10.0 * np.log10(255.0 ** 2 / mse))


def msssim_x_vs_rec(x, x_rec):
    """The reference's literal call (utils.py:94-99): HWC arrays expanded to (H,W,3,1)."""
    a = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()[None]
    b = torch.as_tensor(np.ascontiguousarray(x_rec, dtype=np.float32)).cuda()[None]
    return np.float32(ops.msssim(a, b, form="reference_call")[0])


def msssim_standard(x, x_rec):
    """Standard MS-SSIM of one HWC image pair ((1,H,W,3) batch)."""
    a = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()[None]
    b = torch.as_tensor(np.ascontiguousarray(x_rec, dtype=np.float32)).cuda()[None]
    return np.float32(ops.msssim(a, b, form="standard")[0])


def save_test_imgs_fn(root_save_img, model_name, x_with_si, i, bpp):
    """PNG writer with the reference's naming and uint8 truncation (utils.py:102-111)."""
    from PIL import Image
    path = os.path.join(root_save_img, model_name)
    os.makedirs(path, exist_ok=True)
    img = Image.fromarray(np.transpose(x_with_si, (1, 2, 0)).astype("uint8"), "RGB")
    img.save(os.path.join(path, str(i) + "_" + "{:.5f}bpp.png".format(bpp)))


def pearson_per_patch(x, y, patch_h=20, patch_w=24):
    """Average Pearson correlation between aligned 20x24x3 patches of two HWC images (utils.py:161-180)."""
    import scipy.stats
    H, W = (x.shape[0] // patch_h) * patch_h, (x.shape[1] // patch_w) * patch_w

    def tiles(a):
        a = np.asarray(a)[:H, :W]
        return a.reshape(H // patch_h, patch_h, W // patch_w, patch_w, 3).transpose(0, 2, 1, 3, 4).reshape(
            (H // patch_h) * (W // patch_w), -1)
    px, py = tiles(x), tiles(y)
    tot = 0.0
    for i in range(px.shape[0]):
        tot += scipy.stats.pearsonr(px[i], py[i])[0]
    return tot / px.shape[0]


def loss_list_saver(x, y, x_rec, y_syn, batch_size, model_name, bpp, root_save_img, msssim_fn=None):
    """Appends per-image bpp / L1 / PSNR / MS-SSIM / MSE(x, y_syn) / average patch Pearson(x, y_syn) to the six
    txt lists of the reference (utils.py:114-158); inputs are NCHW batches.  msssim_fn defaults to the CUDA
    MS-SSIM kernel through `msssim_x_vs_rec`."""
    msssim_fn = msssim_fn or msssim_x_vs_rec
    x, y, x_rec, y_syn = (np.transpose(a, (0, 2, 3, 1)) for a in (x, y, x_rec, y_syn))
    names = ["bpp_list_", "l1_list_", "psnr_list_", "msssim_list_", "mse_list_x_y_syn_", "avg_Pearson_list_x_y_syn_"]
    files = [open(root_save_img + n + str(model_name) + ".txt", "a+") for n in names]
    try:
        for i in range(batch_size):
            files[0].write(str(bpp) + "\n")
            files[1].write(str(l1_x_vs_rec(x[i], x_rec[i])[1]) + "\n")
            files[2].write(str(psnr_x_vs_rec(x[i], x_rec[i])) + "\n")
            files[3].write(str(msssim_fn(x[i], x_rec[i])) + "\n")
            files[4].write(str(np.mean((x[i].astype("float32") - y_syn[i].astype("float32")) ** 2)) + "\n")
            files[5].write(str(pearson_per_patch(x[i], y_syn[i])) + "\n")
    finally:
        for f in files:
            f.close()
