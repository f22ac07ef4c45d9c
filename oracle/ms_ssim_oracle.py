"""MS-SSIM oracle (numpy/scipy, float64).  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/ms_ssim_np_imgcomp.py:43-200 (MultiScaleSSIM,
_SSIMForMultiScale, _FSpecialGauss) and the way /root/reference/src/utils.py:94-99 calls
it.  PINNED: tests/golden/msssim_golden.npz holds outputs of the reference module itself
(run here with a stub ``tensorflow`` import) for seeded inputs; tests/test_oracle_golden.py
checks this restatement against them.
"""
import numpy as np
from scipy import signal
from scipy.ndimage import convolve

WEIGHTS = np.array([0.0448, 0.2856, 0.3001, 0.2363, 0.1333])  # ms_ssim_np_imgcomp.py:91-92


def gauss_window(size, sigma):
    """fspecial('gaussian') (ms_ssim_np_imgcomp.py:113-124)."""
    radius = size // 2
    offset, start, stop = 0.0, -radius, radius + 1
    if size % 2 == 0:
        offset, stop = 0.5, stop - 1
    x, y = np.mgrid[offset + start:stop, offset + start:stop]
    g = np.exp(-((x ** 2 + y ** 2) / (2.0 * sigma ** 2)))
    return g / g.sum()


def ssim_cs(img1, img2, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """Mean SSIM and contrast-structure of one scale (ms_ssim_np_imgcomp.py:127-200)."""
    img1 = img1.astype(np.float64)
    img2 = img2.astype(np.float64)
    _, height, width, _ = img1.shape
    size = min(filter_size, height, width)
    sigma = size * filter_sigma / filter_size if filter_size else 0
    if filter_size:
        window = np.reshape(gauss_window(size, sigma), (1, size, size, 1))
        mu1 = signal.fftconvolve(img1, window, mode="valid")
        mu2 = signal.fftconvolve(img2, window, mode="valid")
        s11 = signal.fftconvolve(img1 * img1, window, mode="valid")
        s22 = signal.fftconvolve(img2 * img2, window, mode="valid")
        s12 = signal.fftconvolve(img1 * img2, window, mode="valid")
    else:
        mu1, mu2, s11, s22, s12 = img1, img2, img1 * img1, img2 * img2, img1 * img2
    mu11, mu22, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s11, s22, s12 = s11 - mu11, s22 - mu22, s12 - mu12
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    v1 = 2.0 * s12 + c2
    v2 = s11 + s22 + c2
    ssim = np.mean(((2.0 * mu12 + c1) * v1) / ((mu11 + mu22 + c1) * v2))
    return ssim, np.mean(v1 / v2)


def multi_scale_ssim(img1, img2, max_val=255):
    """MultiScaleSSIM on (batch, height, width, depth) arrays (ms_ssim_np_imgcomp.py:51-110)."""
    if img1.shape != img2.shape or img1.ndim != 4:
        raise RuntimeError("expected two equal-shape 4-D arrays")
    levels = WEIGHTS.size
    down = np.ones((1, 2, 2, 1)) / 4.0
    im1, im2 = img1.astype(np.float64), img2.astype(np.float64)
    mssim, mcs = [], []
    for _ in range(levels):
        s, c = ssim_cs(im1, im2, max_val=max_val)
        mssim.append(s)
        mcs.append(c)
        im1, im2 = [convolve(im, down, mode="reflect")[:, ::2, ::2, :] for im in (im1, im2)]
    mssim, mcs = np.array(mssim), np.array(mcs)
    return np.prod(mcs[:levels - 1] ** WEIGHTS[:levels - 1]) * (mssim[levels - 1] ** WEIGHTS[levels - 1])


def msssim_standard(x_hwc, rec_hwc):
    """Standard form: one image, (1,H,W,3)."""
    return np.float32(multi_scale_ssim(x_hwc[None], rec_hwc[None]))


def msssim_reference_call(x_hwc, rec_hwc):
    """The reference's literal call (utils.py:94-99): HWC expanded to (H,W,3,1) and so read
    as batch=H, height=W, width=3."""
    return np.float32(multi_scale_ssim(x_hwc[..., None], rec_hwc[..., None]))
