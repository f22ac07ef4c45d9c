"""SI_full_img (/root/reference/src/siFull_img.py:5-42) on libdsin_b200.

The reference loops over images and tiles/un-tiles patches with extract_image_patches and a
gradient trick (:45-68); here the whole batch goes through one prepare/match/gather sequence.
Returns the reference's 9-tuple; ``ncc``, ``x_patches`` and ``y_patches`` are None (never
materialised)."""
from __future__ import annotations

from . import ops
from .siFinder import GaussianPrior, match_images


def SI_full_img(x_dec, y_imgs, mask, patch_h, patch_w, ae_config, y_dec):
    x_nhwc = getattr(x_dec, "_dsin_nhwc", None)
    if x_nhwc is None:
        x_nhwc = ops.nchw_to_nhwc(x_dec.contiguous())
    yd_nhwc = getattr(y_dec, "_dsin_nhwc", None)
    if yd_nhwc is None:
        yd_nhwc = ops.nchw_to_nhwc(y_dec.contiguous())
    y_nhwc = getattr(y_imgs, "_dsin_nhwc", None)
    if y_nhwc is None:
        y_nhwc = ops.nchw_to_nhwc(y_imgs.contiguous())
    use_mask = isinstance(mask, GaussianPrior)
    y_syn_nhwc, q, r, row, col, best = match_images(x_nhwc, y_nhwc, yd_nhwc, patch_h, patch_w, use_mask, ae_config)
    y_syn = ops.nhwc_to_nchw(y_syn_nhwc)
    y_syn._dsin_nhwc = y_syn_nhwc
    y_syn._dsin_best = best
    ncc_w = x_dec.shape[3] - patch_w + 1
    return y_syn, None, row * ncc_w + col, q, r, row, col, None, None
