"""SI-Finder edge cases against the oracle (src/siFinder.py:13-33,87-133; Eigen argmax: NaN never wins, all-NaN -> 0,
first index on ties): degenerate windows, exact ties from periodic content, more near-ties than a work unit's
candidate list holds, identical x and y (the prior's peak sits at origin + 1)."""
import numpy as np
import pytest
import torch

from oracle import dsin_oracle as O

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _match(x, y_dec, y=None, use_mask=True):
    from dsin_b200.siFinder import match_images
    y = y_dec if y is None else y
    xt, yt, ydt = (torch.as_tensor(a, dtype=torch.float32) for a in (x, y, y_dec))
    ref_syn, ref_row, ref_col, ref_best = O.si_full_img(xt, yt, ydt, use_mask=use_mask)
    y_syn, _q, _r, row, col, best = match_images(_nhwc(xt.cuda()), _nhwc(yt.cuda()), _nhwc(ydt.cuda()), 20, 24, use_mask)
    return (row.cpu(), col.cpu(), best.cpu(), y_syn.permute(0, 3, 1, 2).cpu()), (ref_row, ref_col, ref_best, ref_syn)


def _near_ties_only(x, y_dec, got, ref, tol, use_mask=True):
    """Every differing (row, col) must be a float64 near-tie of the masked score; returns the number of differences."""
    row, col = got[0], got[1]
    mism = (row != ref[0]) | (col != ref[1])
    N, _, H, W = x.shape
    mask = O.gaussian_masks(H, W, 20, 24) if use_mask else None
    for n, p in zip(*np.nonzero(mism.numpy())):
        xi = torch.as_tensor(x[n], dtype=torch.float64).permute(1, 2, 0)
        yi = torch.as_tensor(y_dec[n], dtype=torch.float64).permute(1, 2, 0)
        q64 = O.rgb_transform(O.sif_normalize_nhwc(O.extract_patches(xi, 20, 24)))
        r64 = O.rgb_transform(O.sif_normalize_nhwc(yi))
        a = O.score_at(q64, r64, mask, p, int(row[n, p]), int(col[n, p]))
        b = O.score_at(q64, r64, mask, p, int(ref[0][n, p]), int(ref[1][n, p]))
        assert abs(a - b) < tol, (n, p, a, b)
    return int(mism.sum())


def _textured(H, W, seed):
    from dsin_b200 import synth
    x, y = synth.make_pair(seed, H, W, sigma=2.0, disparity=9, noise=3.0)
    return x[None], y[None]


def test_all_nan_side_image_gives_index_zero():
    """y_dec equal to the SI-Finder's channel means everywhere: the normalised image is exactly 0, every Pearson value
    is 0/0 = NaN, and tf.argmax over all-NaN returns index 0 -> (row, col) = (0, 0) for every patch."""
    x, _ = _textured(80, 144, 3)
    y_dec = np.broadcast_to(O.SIF_MEAN.reshape(1, 3, 1, 1), x.shape).astype(np.float32).copy()
    got, ref = _match(x, y_dec)
    assert int(ref[0].abs().sum()) == 0 and int(ref[1].abs().sum()) == 0  # the oracle's own statement of the rule
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    assert torch.equal(got[3], ref[3])                                     # gather at (0, 0)


def test_nan_windows_never_win():
    """Left part of y_dec is the NaN-producing constant, the rest is textured: NaN windows are skipped."""
    x, y = _textured(80, 144, 5)
    y_dec = y.copy()
    y_dec[:, :, :, :40] = O.SIF_MEAN.reshape(1, 3, 1, 1)
    got, ref = _match(x, y_dec, y=y)
    n = _near_ties_only(x, y_dec, got, ref, 3e-6)
    assert n <= 1
    assert bool((ref[1] >= 17).all())  # no window that lies wholly inside the constant region was chosen
    assert bool(torch.isfinite(got[2]).all())


def test_flat_patch_with_negative_variance_is_all_nan():
    """A patch equal to the channel means has den_x = 0 or a rounding-sized negative number: every score is NaN or the
    patch is skipped by the same rule -> index 0, as in the oracle."""
    x, y = _textured(80, 144, 7)
    x[0, :, 20:40, 24:48] = O.SIF_MEAN.reshape(3, 1, 1)  # patch p = 7
    got, ref = _match(x, y)
    assert int(ref[0][0, 7]) == 0 and int(ref[1][0, 7]) == 0
    n = _near_ties_only(x, y, got, ref, 3e-6)
    assert n <= 1 and int(got[0][0, 7]) == 0 and int(got[1][0, 7]) == 0


def test_identical_images_match_in_place():
    """y = x: every patch correlates perfectly with its own location; the prior's peak is at (top + 1, left + 1)
    (src/AE.py:193-220), the exact match at (top, left) still wins."""
    x, _ = _textured(80, 144, 9)
    got, ref = _match(x, x.copy())
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    P = ref[0].shape[1]
    top = torch.tensor([(p // 6) * 20 for p in range(P)], dtype=torch.int32)
    left = torch.tensor([(p % 6) * 24 for p in range(P)], dtype=torch.int32)
    assert torch.equal(ref[0][0], top) and torch.equal(ref[1][0], left)
    assert torch.equal(got[3], ref[3])


def test_exact_ties_resolve_to_the_first_index():
    """Periodic content and no prior: all periods score bit-identically; tf.argmax takes the first index."""
    H, W = 80, 144
    yy, xx = np.mgrid[0:H, 0:W]
    base = 128 + 64 * np.sin(2 * np.pi * xx / 8.0) * np.cos(2 * np.pi * yy / 10.0)
    img = np.stack([base, 0.5 * base + 40, 255 - base], 0)[None].astype(np.float32)
    img = np.floor(img)
    got, ref = _match(img, img.copy(), use_mask=False)
    n = _near_ties_only(img, img, got, ref, 3e-6, use_mask=False)
    print("periodic, no prior: %d / %d differ (float64 near-ties)" % (n, ref[0].numel()))
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])


def test_more_near_ties_than_a_work_unit_keeps():
    """320x1224, horizontally periodic texture (period 8) under the prior: around the prior's peak several periods lie
    within the coarse scorer's tolerance of each other inside ONE work unit (4 rows x 640 columns, 4 candidates kept),
    so the exact winner must come from the exhaustive group rescoring."""
    H, W = 320, 1224
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:H, 0:W]
    col_code = rng.uniform(0, 255, size=(3, H, 8))               # a random 8-periodic pattern per row and channel
    img = col_code[:, :, xx[0] % 8][None].astype(np.float32)     # (1,3,H,W)
    img = np.floor(img)
    got, ref = _match(img, img.copy())
    n = _near_ties_only(img, img, got, ref, 3e-6)
    print("periodic under the prior: %d / %d differ (float64 near-ties)" % (n, ref[0].numel()))
    assert n <= 4
