"""CPU oracle for the DSIN inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (torch-CPU / numpy, fp32 by default, fp64 on request) of
the arithmetic of the reference TF1 graph (ayziksha/DSIN @ a3b8d05).  It is imported
only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- never by the product package ``dsin_b200``.

PARITY UNPINNED: the reference ships no tests, golden vectors or
fixtures, and TensorFlow 1.11 cannot be installed here, so the TF op semantics
restated below (SAME padding, FusedBatchNorm inference, Conv3D, ExtractImagePatches,
CropAndResize, ArgMax tie-breaking, softmax cross entropy) are pinned only by the TF
documentation/kernels as summarised in SURVEY.md App. A.  The two pieces of the
reference that *do* run here -- ``AE.create_gaussian_masks`` (pure numpy), ``ms_ssim_np_imgcomp``
(numpy/scipy) and the numpy-only pieces of the model code (causal conv masks, symbol-volume padding,
``AE.normalize``/``denormalize``/``get_mean_var``, the coder helpers' block order) -- are pinned by
``tests/golden`` fixtures that were generated from the reference itself (``tests/golden/make_golden.py``).

All citations are ``file:line`` relative to /root/reference/.
Weights are a flat ``dict[str, np.ndarray]`` keyed by the TF variable names of
SURVEY.md App. A.11 (conv weights HWIO, transposed-conv [k,k,out,in], conv3d DHWIO).
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

# src/autoencoder_imgcomp.py:15
EncoderOutput = namedtuple("EncoderOutput", ["qbar", "qhard", "symbols", "z", "heatmap"])

ENC = "encoder/encoder_body/encoder_body/autoencoder/encoder/"
DEC = "decoder/autoencoder/decoder/"
PC = "imgcomp/probclass3d/logits/"
SIN = "siNetwork/"

# src/autoencoder_imgcomp.py:160-170, src/AE.py:240-250 (float32 constants)
KITTI_MEAN = np.array([93.70454143384742, 98.28243432206516, 94.84678088809876], dtype=np.float32)
KITTI_VAR = np.array([5411.79935676, 5758.60456747, 5890.31451232], dtype=np.float32)
# src/siFinder.py:62-64 (python floats -> float32 constants; named "variances", used as divisors)
SIF_MEAN = np.array([93.70454143384742, 98.28243432206516, 94.84678088809876], dtype=np.float32)
SIF_DIV = np.array([73.56493292844912, 75.88547006820752, 76.74838442810665], dtype=np.float32)

BN_EPS = 1e-5  # src/autoencoder_imgcomp.py:119
HARD_SIGMA = 1e7  # src/quantizer_imgcomp.py:5


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


# --------------------------------------------------------------------------------------
# normalisation  (src/autoencoder_imgcomp.py:136-154, src/AE.py:222-238)
# --------------------------------------------------------------------------------------
def _std(dtype):
    # np.sqrt(var + 1e-10) is evaluated in float32 by numpy (float32 array + python float)
    return _t(np.sqrt(KITTI_VAR + 1e-10), dtype).view(1, 3, 1, 1)


def normalize(x):
    return (x - _t(KITTI_MEAN, x.dtype).view(1, 3, 1, 1)) / _std(x.dtype)


def denormalize(x):
    return x * _std(x.dtype) + _t(KITTI_MEAN, x.dtype).view(1, 3, 1, 1)


# --------------------------------------------------------------------------------------
# TF SAME padding and slim conv blocks (src/autoencoder_imgcomp.py:106-125)
# --------------------------------------------------------------------------------------
def same_pads(n, k, s, d=1):
    out = -(-n // s)
    keff = (k - 1) * d + 1
    total = max((out - 1) * s + keff - n, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w_hwio, stride=1, dilation=1):
    """tf.nn.conv2d(..., padding='SAME') on NCHW input with an HWIO filter."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = same_pads(x.shape[2], kh, stride, dilation)
    pl, pr = same_pads(x.shape[3], kw, stride, dilation)
    w = _t(w_hwio, x.dtype).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, stride=stride, dilation=dilation)


def conv2d_transpose_same_s2(x, w_kkoi):
    """slim.conv2d_transpose(stride=2, SAME) with filter [k,k,out,in]: the adjoint of the
    SAME stride-2 forward conv = full transposed conv cropped at [before : before+2n]
    (src/autoencoder_imgcomp.py:252,265,266)."""
    k = w_kkoi.shape[0]
    n_h, n_w = x.shape[2], x.shape[3]
    bh = same_pads(2 * n_h, k, 2)[0]
    bw = same_pads(2 * n_w, k, 2)[0]
    w = _t(w_kkoi, x.dtype).permute(3, 2, 0, 1).contiguous()  # (in, out, k, k)
    full = F.conv_transpose2d(x, w, stride=2)
    return full[:, :, bh:bh + 2 * n_h, bw:bw + 2 * n_w]


def batch_norm_inference(x, W, scope):
    """FusedBatchNorm, is_training=False: (x-mean) * (gamma*rsqrt(var+eps)) + beta."""
    dt = x.dtype
    g = _t(W[scope + "/BatchNorm/gamma"], dt)
    b = _t(W[scope + "/BatchNorm/beta"], dt)
    m = _t(W[scope + "/BatchNorm/moving_mean"], dt)
    v = _t(W[scope + "/BatchNorm/moving_variance"], dt)
    s = g * torch.rsqrt(v + BN_EPS)
    return (x - m.view(1, -1, 1, 1)) * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def conv_block(x, W, scope, stride=1, relu=True, trace=None):
    y = conv2d_same(x, W[scope + "/weights"], stride=stride)
    if trace is not None:
        trace(scope, y)
    y = batch_norm_inference(y, W, scope)
    return torch.relu(y) if relu else y


def deconv_block(x, W, scope, relu=True, trace=None):
    y = conv2d_transpose_same_s2(x, W[scope + "/weights"])
    if trace is not None:
        trace(scope, y)
    y = batch_norm_inference(y, W, scope)
    return torch.relu(y) if relu else y


def residual_block(x, W, scope, relu_first=True, trace=None):
    """src/autoencoder_imgcomp.py:275-288.  activation_fn=None is forwarded to BOTH convs
    for res_block_enc_final / dec_after_res."""
    y = conv_block(x, W, scope + "/conv1", relu=relu_first, trace=trace)
    y = conv_block(y, W, scope + "/conv2", relu=False, trace=trace)
    return y + x


# --------------------------------------------------------------------------------------
# heatmap + quantiser (src/autoencoder_imgcomp.py:173-201, src/quantizer_imgcomp.py:43-100)
# --------------------------------------------------------------------------------------
def heatmap3d(z33):
    C = z33.shape[1] - 1
    h = torch.sigmoid(z33[:, 0]) * C
    c = torch.arange(C, dtype=z33.dtype).view(1, C, 1, 1)
    return torch.clamp(torch.clamp(h.unsqueeze(1) - c, max=1.0), min=0.0)


def quantize(z, centers):
    c = _t(centers, z.dtype)
    d = torch.square(torch.abs(z.unsqueeze(-1) - c))
    phi_soft = torch.softmax(-1.0 * d, dim=-1)
    phi_hard = torch.softmax(-HARD_SIGMA * d, dim=-1)
    symbols = torch.argmax(phi_hard, dim=-1)  # first maximal index, int64
    qsoft = (phi_soft * c).sum(-1)
    qhard = c[symbols]
    qbar = qsoft + (qhard - qsoft)  # src/autoencoder_imgcomp.py:132-133
    return qbar, qsoft, qhard, symbols


# --------------------------------------------------------------------------------------
# encoder / decoder (src/autoencoder_imgcomp.py:219-269)
# --------------------------------------------------------------------------------------
def encode(x, W, B=5, trace=None, force_symbols=None):
    """force_symbols (test helper): int tensor (N,C,h,w); where it is >= 0 the quantiser's decision is overridden
    by it (qhard, qbar follow).  Used to replay the oracle downstream of an adjudicated near-tie."""
    net = normalize(x)
    net = conv_block(net, W, ENC + "h1", stride=2, trace=trace)
    net = conv_block(net, W, ENC + "h2", stride=2, trace=trace)
    r0 = net
    for b in range(B):
        rb = net
        for i in (1, 2, 3):
            net = residual_block(net, W, ENC + "res_block_enc_%d/enc_%d_%d" % (b, b, i), trace=trace)
        net = net + rb
    net = residual_block(net, W, ENC + "res_block_enc_final", relu_first=False, trace=trace)
    net = net + r0
    z33 = conv_block(net, W, ENC + "to_bn", stride=2, relu=False, trace=trace)
    hm = heatmap3d(z33)
    z = hm * z33[:, 1:]
    qbar, qsoft, qhard, symbols = quantize(z, W[ENC + "centers"])
    if force_symbols is not None:
        fs = torch.as_tensor(force_symbols).to(torch.int64)
        symbols = torch.where(fs >= 0, fs, symbols)
        qhard = _t(W[ENC + "centers"], z.dtype)[symbols]
        qbar = qsoft + (qhard - qsoft)
    return EncoderOutput(qbar, qhard, symbols, z, hm)


def decode(q, W, B=5, trace=None):
    net = deconv_block(q, W, DEC + "from_bn", trace=trace)
    r0 = net
    for b in range(B):
        rb = net
        for i in (1, 2, 3):
            net = residual_block(net, W, DEC + "res_block_dec_%d/dec_%d_%d" % (b, b, i), trace=trace)
        net = net + rb
    net = residual_block(net, W, DEC + "dec_after_res", relu_first=False, trace=trace)
    net = net + r0
    net = deconv_block(net, W, DEC + "h12", trace=trace)
    net = deconv_block(net, W, DEC + "h13", relu=False, trace=trace)
    net = denormalize(net)
    return torch.clamp(net, 0.0, 255.0)


# --------------------------------------------------------------------------------------
# probability classifier (src/probclass_imgcomp.py:63-106,150-196,214-261,268-292)
# --------------------------------------------------------------------------------------
def pc_masks(K=3):
    first = np.ones((K // 2 + 1, K, K), dtype=np.float32)
    first[-1, K // 2, K // 2:] = 0
    first[-1, K // 2 + 1:, :] = 0
    other = np.ones((K // 2 + 1, K, K), dtype=np.float32)
    other[-1, K // 2, K // 2 + 1:] = 0
    other[-1, K // 2 + 1:, :] = 0
    return first, other


def _conv3d(x, W, scope, mask, relu):
    dt = x.dtype
    w = _t(W[scope + "/weights"], dt) * _t(mask, dt)[..., None, None]  # DHWio * DHW11
    w = w.permute(4, 3, 0, 1, 2).contiguous()  # (out, in, D, H, W)
    y = F.conv3d(x, w, bias=_t(W[scope + "/biases"], dt))
    return torch.relu(y) if relu else y


def probclass_bitcost(qbar, symbols, W, num_centers=6):
    """-> bits per symbol, (N,C,H,W).  The bottleneck channel axis is the conv depth."""
    dt = qbar.dtype
    first, other = pc_masks(3)
    pad_value = float(np.asarray(W[ENC + "centers"], dtype=np.float32)[0])  # auto_pad_value, :59-61
    pad = 4  # context_size 9 // 2
    x = F.pad(qbar, (pad, pad, pad, pad, pad, 0), value=pad_value)  # C:[4,0] H:[4,4] W:[4,4]
    x = x.unsqueeze(1)  # (N, 1, D=C+4, H+8, W+8)
    net = _conv3d(x, W, PC + "conv3d_conv0_mask", first, True)
    res_in = net
    net = _conv3d(net, W, PC + "res1/conv3d_conv1_mask", other, True)
    net = _conv3d(net, W, PC + "res1/conv3d_conv2_mask", other, False)
    net = net + res_in[:, :, 2:, 2:-2, 2:-2]
    logits = _conv3d(net, W, PC + "conv3d_conv2_mask", other, True)  # default activation = relu (F10)
    # softmax_cross_entropy_with_logits * log2(e)
    lse = torch.logsumexp(logits, dim=1)
    picked = torch.gather(logits, 1, symbols.unsqueeze(1)).squeeze(1)
    log2e = torch.tensor(np.float32(np.log2(np.e))).to(dt)
    return (lse - picked) * log2e


def bitcost_to_bpp(bitcost, n_pixels_total):
    """src/bits_imgcomp.py:4-20: sum(bits) / (N*H*W)."""
    return bitcost.sum() / float(n_pixels_total)


# --------------------------------------------------------------------------------------
# SI-Finder (src/AE.py:193-220, src/siFull_img.py:5-68, src/siFinder.py:7-135)
# --------------------------------------------------------------------------------------
def gaussian_masks(H, W, ph, pw):
    """Restatement of AE.create_gaussian_masks (src/AE.py:193-220): float64 math, cast to
    float32, returned as (P, H-ph+1, W-pw+1) (reference layout is (1, h, w, P))."""
    n = np.arange(0, (H * W) // (ph * pw))
    patch_img_w = W / pw
    w = np.arange(0, W, 1, float)
    h = np.arange(0, H, 1, float)
    ch = (n // patch_img_w + 0.5) * ph
    cw = ((n % patch_img_w) + 0.5) * pw
    sh, sw = 0.5 * H, 0.5 * W
    cols = (w[None, :] - cw[:, None]) ** 2 / sw ** 2  # (P, W)
    rows = (h[None, :] - ch[:, None]) ** 2 / sh ** 2  # (P, H)
    g = np.exp(-4 * np.log(2) * (rows[:, :, None] + cols[:, None, :]))
    g = g[:, ph // 2 - 1:H - ph // 2, pw // 2 - 1:W - pw // 2]
    return g.astype(np.float32)


def sif_normalize_nhwc(img_nhwc):
    dt = img_nhwc.dtype
    return (img_nhwc - _t(SIF_MEAN, dt)) / _t(SIF_DIV, dt)


def rgb_transform(x_nhwc):
    R, G, B = x_nhwc[..., 0:1], x_nhwc[..., 1:2], x_nhwc[..., 2:3]
    return torch.cat([R + G, R - G, 0.5 * (R + B)], dim=-1)  # src/siFinder.py:149-153


def extract_patches(img_hwc, ph, pw):
    """tf.extract_image_patches, ksize=stride=(ph,pw): (P, ph, pw, C), p = pr*(W/pw)+pc."""
    H, W, C = img_hwc.shape
    t = img_hwc.reshape(H // ph, ph, W // pw, pw, C).permute(0, 2, 1, 3, 4)
    return t.reshape(-1, ph, pw, C)


def fold_patches(patches, H, W):
    P, ph, pw, C = patches.shape
    t = patches.reshape(H // ph, W // pw, ph, pw, C).permute(0, 2, 1, 3, 4)
    return t.reshape(H, W, C)


def pearson_scores(q, r_hwc, mask, chunk=64):
    """Masked Pearson score per (patch, position) and its argmax.

    q: (P,ph,pw,3) transformed patches, r_hwc: (H,W,3) transformed search image,
    mask: (P,h,w) float32 numpy.  Returns idx (P,) int64 [first maximal index, NaN never
    wins], best score (P,), and optionally nothing else (the 1.18 GB map is chunked).
    src/siFinder.py:87-133 evaluated literally, left to right, in q.dtype."""
    dt = q.dtype
    P, ph, pw, C = q.shape
    n = float(ph * pw * C)
    r = r_hwc.permute(2, 0, 1).unsqueeze(0)  # (1,3,H,W)
    ones = torch.ones(1, C, ph, pw, dtype=dt)
    sum_y = F.conv2d(r, ones)[0, 0]
    sum_y2 = F.conv2d(r * r, ones)[0, 0]
    y_mean = F.conv2d(r, torch.full((1, C, ph, pw), 1.0 / n, dtype=dt))[0, 0]
    qf = q.reshape(P, -1)
    sum_x = qf.sum(1)
    sum_x2 = (qf * qf).sum(1)
    x_mean = qf.mean(1)
    den_y = sum_y2 - 2 * (y_mean * sum_y) + n * (y_mean * y_mean)
    den_x = sum_x2 - 2 * (x_mean * sum_x) + n * (x_mean * x_mean)
    idx_out = torch.zeros(P, dtype=torch.int64)
    best_out = torch.zeros(P, dtype=dt)
    filt = q.permute(0, 3, 1, 2).contiguous()  # (P,3,ph,pw)
    for s in range(0, P, chunk):
        e = min(P, s + chunk)
        xy = F.conv2d(r, filt[s:e])[0]  # (p,h,w)
        sx = sum_x[s:e].view(-1, 1, 1)
        xm = x_mean[s:e].view(-1, 1, 1)
        num = xy - y_mean * sx - sum_y * xm + n * (y_mean * xm)
        den = den_y * den_x[s:e].view(-1, 1, 1)
        ncc = num / torch.sqrt(den)
        if mask is not None:
            ncc = ncc * _t(mask[s:e], dt)
        flat = ncc.reshape(e - s, -1)
        # tf.argmax (Eigen): NaN never beats a number; all-NaN -> 0; first index on ties
        clean = torch.where(torch.isnan(flat), torch.full_like(flat, -float("inf")), flat)
        idx = torch.argmax(clean, dim=1)
        allnan = torch.isnan(flat).all(dim=1)
        idx = torch.where(allnan, torch.zeros_like(idx), idx)
        idx_out[s:e] = idx
        best_out[s:e] = flat.gather(1, idx.view(-1, 1)).view(-1)
    return idx_out, best_out


def score_at(q, r_hwc, mask, p, row, col):
    """Masked Pearson score of patch p at one position, same formula (used by tests to
    adjudicate near-ties in float64)."""
    dt = q.dtype
    ph, pw, C = q.shape[1:]
    n = float(ph * pw * C)
    win = r_hwc[row:row + ph, col:col + pw, :]
    qp = q[p]
    xy = (qp * win).sum()
    sy, sy2, ym = win.sum(), (win * win).sum(), win.sum() / n
    sx, sx2, xm = qp.sum(), (qp * qp).sum(), qp.mean()
    num = xy - ym * sx - sy * xm + n * (ym * xm)
    den = (sy2 - 2 * ym * sy + n * ym * ym) * (sx2 - 2 * xm * sx + n * xm * xm)
    v = num / torch.sqrt(den)
    if mask is not None:
        v = v * float(mask[p, row, col])
    return float(v)


def crop_and_resize_patches(y_hwc, rows, cols, ph, pw):
    """tf.image.crop_and_resize with the reference's boxes (src/siFinder.py:35-41;
    SURVEY App. A.9), all coordinate arithmetic in float32.  y_hwc numpy float32."""
    H, W, C = y_hwc.shape
    f32 = np.float32
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    y1 = (rows.astype(np.float64) / H).astype(f32)
    x1 = (cols.astype(np.float64) / W).astype(f32)
    y2 = ((rows + ph).astype(np.float64) / H).astype(f32)
    x2 = ((cols + pw).astype(np.float64) / W).astype(f32)
    P = rows.shape[0]
    out = np.zeros((P, ph, pw, C), dtype=f32)
    hs = ((y2 - y1) * f32(H - 1) / f32(ph - 1)).astype(f32)
    ws = ((x2 - x1) * f32(W - 1) / f32(pw - 1)).astype(f32)
    ty = np.arange(ph, dtype=f32)
    tx = np.arange(pw, dtype=f32)
    in_y = (y1[:, None] * f32(H - 1) + ty[None, :] * hs[:, None]).astype(f32)  # (P,ph)
    in_x = (x1[:, None] * f32(W - 1) + tx[None, :] * ws[:, None]).astype(f32)  # (P,pw)
    for p in range(P):
        iy, ix = in_y[p], in_x[p]
        vy = ~((iy < 0) | (iy > f32(H - 1)))
        vx = ~((ix < 0) | (ix > f32(W - 1)))
        top = np.floor(iy).astype(np.int64).clip(0, H - 1)
        bot = np.ceil(iy).astype(np.int64).clip(0, H - 1)
        ly = (iy - np.floor(iy)).astype(f32)
        lef = np.floor(ix).astype(np.int64).clip(0, W - 1)
        rig = np.ceil(ix).astype(np.int64).clip(0, W - 1)
        lx = (ix - np.floor(ix)).astype(f32)
        tl = y_hwc[top][:, lef]  # (ph,pw,C)
        tr = y_hwc[top][:, rig]
        bl = y_hwc[bot][:, lef]
        br = y_hwc[bot][:, rig]
        lxb = lx[None, :, None]
        T = (tl + (tr - tl) * lxb).astype(f32)
        Bm = (bl + (br - bl) * lxb).astype(f32)
        v = (T + (Bm - T) * ly[:, None, None]).astype(f32)
        v[~vy, :, :] = 0
        v[:, ~vx, :] = 0
        out[p] = v
    return out


def si_finder(x_patches, y_hwc, mask, ph, pw, y_dec_hwc, force_rowcol=None):
    """src/siFinder.py:7-53 (Pearson branch, batch_size == 1).  force_rowcol (test helper): (row, col) int arrays,
    entries >= 0 override the argmax (replay downstream of an adjudicated near-tie)."""
    q = rgb_transform(sif_normalize_nhwc(x_patches))
    r = rgb_transform(sif_normalize_nhwc(y_dec_hwc))
    idx, best = pearson_scores(q, r, mask)
    ncc_w = y_dec_hwc.shape[1] - pw + 1
    row = (idx // ncc_w).to(torch.int32)
    col = (idx % ncc_w).to(torch.int32)
    if force_rowcol is not None:
        fr, fc = (torch.as_tensor(np.asarray(a)).to(torch.int32) for a in force_rowcol)
        row = torch.where(fr >= 0, fr, row)
        col = torch.where(fc >= 0, fc, col)
    y_np = y_hwc.to(torch.float32).numpy()
    yp = crop_and_resize_patches(y_np, row.numpy(), col.numpy(), ph, pw)
    return torch.from_numpy(yp).to(x_patches.dtype), best, q, r, row, col


def si_full_img(x_dec, y, y_dec, ph=20, pw=24, use_mask=True, force_rowcol=None):
    """src/siFull_img.py:5-42: returns y_syn NCHW plus (row, col, best) per image."""
    N, C, H, W = x_dec.shape
    mask = gaussian_masks(H, W, ph, pw) if use_mask else None
    outs, rows, cols, bests = [], [], [], []
    for n in range(N):
        xi = x_dec[n].permute(1, 2, 0)
        yi = y[n].permute(1, 2, 0)
        ydi = y_dec[n].permute(1, 2, 0)
        xp = extract_patches(xi, ph, pw)
        frc = None if force_rowcol is None else (force_rowcol[0][n], force_rowcol[1][n])
        yp, best, _q, _r, row, col = si_finder(xp, yi, mask, ph, pw, ydi, frc)
        outs.append(fold_patches(yp, H, W).permute(2, 0, 1))
        rows.append(row)
        cols.append(col)
        bests.append(best)
    return torch.stack(outs), torch.stack(rows), torch.stack(cols), torch.stack(bests)


# --------------------------------------------------------------------------------------
# SI-Net (src/siNet.py:9-41, src/AE.py:63-69)
# --------------------------------------------------------------------------------------
SINET_RATES = (1, 2, 4, 8, 16, 32, 64, 128, 1)


def si_net(inp, W):
    net = inp
    for i, rate in enumerate(SINET_RATES):
        sc = SIN + "g_conv%d" % (i + 1)
        net = conv2d_same(net, W[sc + "/weights"], dilation=rate) + _t(W[sc + "/biases"], inp.dtype).view(1, -1, 1, 1)
        net = torch.maximum(net * 0.2, net)
    sc = SIN + "g_conv_last"
    return conv2d_same(net, W[sc + "/weights"]) + _t(W[sc + "/biases"], inp.dtype).view(1, -1, 1, 1)


# --------------------------------------------------------------------------------------
# AE.siNet_get_reconstructed (src/AE.py:132-152)
# --------------------------------------------------------------------------------------
Reconstruction = namedtuple(
    "Reconstruction", ["y_dec", "y_syn", "x_dec", "x_with_si", "bpp", "symbols", "row", "col", "bits_per_image", "best"])


def ae_pass(x, W, force_symbols=None):
    enc = encode(x, W, force_symbols=force_symbols)
    return enc, decode(enc.qbar, W)


def reconstruct(x_np, y_np, W, dtype=torch.float32, ph=20, pw=24, use_mask=True, force_symbols_x=None,
                force_symbols_y=None, force_rowcol=None):
    """x_np, y_np: (N,3,H,W) uint8-valued arrays.  Each pair is processed with the
    reference's batch-1 semantics; bpp is the batch aggregate (src/bits_imgcomp.py:13-14).
    force_*: test helpers that override quantiser / argmax decisions at adjudicated near-ties (entries >= 0)."""
    x = _t(x_np, dtype)
    y = _t(y_np, dtype)
    N, _, H, Wd = x.shape
    with torch.no_grad():
        _ency, y_dec = ae_pass(y, W, force_symbols_y)  # create_y_dec, src/AE.py:150-152
        encx, x_dec = ae_pass(x, W, force_symbols_x)
        bits = probclass_bitcost(encx.qbar, encx.symbols, W)
        bpp = bitcost_to_bpp(bits, N * H * Wd)
        y_syn, row, col, best = si_full_img(x_dec, y, y_dec, ph, pw, use_mask, force_rowcol)
        cat = torch.cat([normalize(x_dec), normalize(y_syn)], dim=1)
        x_with_si = denormalize(si_net(cat, W))
    return Reconstruction(y_dec, y_syn, x_dec, x_with_si, bpp, encx.symbols, row, col,
                          bits.reshape(N, -1).sum(1), best)


# --------------------------------------------------------------------------------------
# validation loss (src/AE.py:71-99,120-131: siNet_validate -> loss_test)
# --------------------------------------------------------------------------------------
ValidationLoss = namedtuple("ValidationLoss", ["loss", "d_loss_scaled", "pc_loss", "reg_loss", "loss_siNet",
                                               "H_real", "H_mask"])


def regularization_loss(W, scope, factor, factor_centers):
    """tf.losses.get_regularization_loss(scope) as the reference calls it (src/autoencoder_imgcomp.py:80-86): the
    REGULARIZATION_LOSSES collection is filtered with re.match(scope, op.name), i.e. a PREFIX match on the full op name.
    slim adds factor * l2_loss(w) = factor * sum(w^2) / 2 for every conv / deconv `weights` variable created inside
    _building_ctx (src/autoencoder_imgcomp.py:98-104) and factor_centers * l2_loss(centers) for the centres
    (src/quantizer_imgcomp.py:18-24); op names start with the variable's full name.  In the graph AE.py builds, those
    names start with 'encoder/encoder_body/...' resp. 'decoder/...' (src/AE.py:51-56) while the scopes asked for are
    'autoencoder/encoder' and 'autoencoder/decoder' (src/autoencoder_imgcomp.py:21-23): nothing matches, the sum is
    0.0.  The rule is restated in full so that a differently scoped weight file behaves as TF would."""
    import re
    rx = re.compile(scope)
    total = 0.0
    for name in sorted(W):
        if not rx.match(name):
            continue
        w = np.asarray(W[name], dtype=np.float64)
        if name.endswith("/weights") and factor:
            total += float(factor) * 0.5 * float(np.sum(w * w))
        elif name.endswith("/centers") and factor_centers:
            total += float(factor_centers) * 0.5 * float(np.sum(w * w))
    return total


def distortion_to_minimize(x, x_out, kind="mae", K_psnr=100.0):
    """Distortions(config, x, x_out, is_training=True).d_loss_scaled (src/Distortions_imgcomp.py:7-56,68-118): per-image
    means, then the batch mean.  With is_training=True the distortion that is minimised is NOT cast to integers (:19-21)."""
    if kind == "mae":
        return torch.abs(x_out - x).mean(dim=(1, 2, 3)).mean()
    if kind == "mse":
        return torch.square(x_out - x).mean(dim=(1, 2, 3)).mean()
    if kind == "psnr":
        mse = torch.square(x_out - x).mean(dim=(1, 2, 3))
        return K_psnr - (10.0 * (torch.log(255.0 * 255.0 / mse) / math.log(10.0))).mean()
    raise ValueError("distortion_to_minimize = %r is not restated (the shipped config uses mae)" % (kind,))


def get_loss(d_loss_scaled, bc, heatmap, beta, H_target, reg_enc=0.0, reg_dec=0.0, reg_probclass=0.0):
    """src/Distortions_imgcomp.py:113-146 -> (total_loss, H_real, H_mask, pc_loss)."""
    H_real = bc.mean()
    H_mask = (bc * heatmap).mean()
    H_soft = 0.5 * (H_mask + H_real)
    pc_loss = beta * torch.clamp(H_soft - H_target, min=0.0)
    reg_loss = reg_probclass + reg_enc + reg_dec
    return d_loss_scaled + pc_loss + reg_loss, H_real, H_mask, pc_loss


def validation_loss(x_np, y_np, W, dtype=torch.float32, ph=20, pw=24, use_mask=True, si_weight=0.7, beta=500.0,
                    H_target=0.04, distortion="mae", K_psnr=100.0, regularization_factor=0.005,
                    regularization_factor_centers=0.1, force_symbols_x=None, force_symbols_y=None, force_rowcol=None):
    """AE.siNet_validate (src/AE.py:120-131): one forward pass in inference mode -> the scalar loss_test =
    (1 - si_weight) * d_loss_scaled + beta * max(0.5 * (mean(bc * heatmap) + mean(bc)) - H_target, 0) + reg
    + si_weight * mean|x - x_with_si|  (src/AE.py:76-99)."""
    x = _t(x_np, dtype)
    y = _t(y_np, dtype)
    with torch.no_grad():
        _ency, y_dec = ae_pass(y, W, force_symbols_y)
        encx, x_dec = ae_pass(x, W, force_symbols_x)
        bc = probclass_bitcost(encx.qbar, encx.symbols, W)
        y_syn, _row, _col, _best = si_full_img(x_dec, y, y_dec, ph, pw, use_mask, force_rowcol)
        x_with_si = denormalize(si_net(torch.cat([normalize(x_dec), normalize(y_syn)], dim=1), W))
        d = distortion_to_minimize(x, x_dec, distortion, K_psnr)
        reg_enc = regularization_loss(W, "autoencoder/encoder", regularization_factor, regularization_factor_centers)
        reg_dec = regularization_loss(W, "autoencoder/decoder", regularization_factor, 0.0)
        total, H_real, H_mask, pc_loss = get_loss((1.0 - si_weight) * d, bc, encx.heatmap, beta, H_target,
                                                  reg_enc, reg_dec)  # probclass: regularization_factor = None
        l_si = torch.abs(x - x_with_si).mean()  # tf.losses.absolute_difference: sum / number of elements
        loss = total + si_weight * l_si
    return ValidationLoss(float(loss), float(d), float(pc_loss), reg_enc + reg_dec, float(l_si), float(H_real),
                          float(H_mask))


# --------------------------------------------------------------------------------------
# BN calibration of random-init weights (SURVEY section 7 hard part 6) -- test helper
# --------------------------------------------------------------------------------------
def calibrate_bn(W, x_np, dtype=torch.float32):
    """One forward pass of encode+decode on x_np that overwrites every BatchNorm's
    moving_mean/moving_variance with the statistics of its own pre-BN input, layer by
    layer, so random-init activations stay O(1).  Mutates and returns W."""
    def trace(scope, y):
        m = y.mean(dim=(0, 2, 3)).double().numpy()
        v = y.var(dim=(0, 2, 3), unbiased=False).double().numpy()
        W[scope + "/BatchNorm/moving_mean"] = m.astype(np.float32)
        W[scope + "/BatchNorm/moving_variance"] = np.maximum(v, 1e-6).astype(np.float32)
    with torch.no_grad():
        enc = encode(_t(x_np, dtype), W, trace=trace)
        decode(enc.qbar, W, trace=trace)
    return W
