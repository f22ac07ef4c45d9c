"""Metric helpers of the reference's utils.py (/root/reference/src/utils.py:82-99) on libdsin_b200.
Plotting and the txt list writers (utils.py:12-79,114-158) are presentation code and out of scope."""
from __future__ import annotations

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
