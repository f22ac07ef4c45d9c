"""Seeded synthetic weights and image pairs.

The reference's pretrained weights are an external download (/root/reference/README.md:38)
and the KITTI data is absent, so benchmarks and parity tests use random-init weights with
the reference's variable names/shapes and structured synthetic stereo
pairs.  Pure numpy: the same dict feeds the CUDA path and the CPU oracle.
"""
from __future__ import annotations

import numpy as np

ENC = "encoder/encoder_body/encoder_body/autoencoder/encoder/"
DEC = "decoder/autoencoder/decoder/"
PC = "imgcomp/probclass3d/logits/"
SIN = "siNetwork/"

SINET_RATES = (1, 2, 4, 8, 16, 32, 64, 128, 1)


def enc_conv_scopes(B=5):
    """Conv scopes of the encoder in execution order: (scope, k, cin, cout)."""
    out = [(ENC + "h1", 5, 3, 64), (ENC + "h2", 5, 64, 128)]
    for b in range(B):
        for i in (1, 2, 3):
            for c in (1, 2):
                out.append((ENC + "res_block_enc_%d/enc_%d_%d/conv%d" % (b, b, i, c), 3, 128, 128))
    for c in (1, 2):
        out.append((ENC + "res_block_enc_final/conv%d" % c, 3, 128, 128))
    return out


def dec_conv_scopes(B=5):
    out = []
    for b in range(B):
        for i in (1, 2, 3):
            for c in (1, 2):
                out.append((DEC + "res_block_dec_%d/dec_%d_%d/conv%d" % (b, b, i, c), 3, 128, 128))
    for c in (1, 2):
        out.append((DEC + "dec_after_res/conv%d" % c, 3, 128, 128))
    return out


def _xavier(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _bn(rng, W, scope, c):
    W[scope + "/BatchNorm/gamma"] = rng.uniform(0.5, 1.5, size=c).astype(np.float32)
    W[scope + "/BatchNorm/beta"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    W[scope + "/BatchNorm/moving_mean"] = np.zeros(c, dtype=np.float32)
    W[scope + "/BatchNorm/moving_variance"] = np.ones(c, dtype=np.float32)


def make_weights(seed=0, B=5, num_chan_bn=32, num_centers=6, pc_k=24, residual_gamma=None):
    """Random-init weights with the reference's names/shapes.  BN moving statistics are
    (0, 1) until ``calibrate`` (oracle- or GPU-driven) overwrites them."""
    rng = np.random.default_rng(seed)
    W = {}
    # quantiser centres: unsorted uniform(-2, 2) draw (src/quantizer_imgcomp.py:28-31)
    W[ENC + "centers"] = rng.uniform(-2, 2, size=num_centers).astype(np.float32)
    for scope, k, cin, cout in enc_conv_scopes(B):
        W[scope + "/weights"] = _xavier(rng, (k, k, cin, cout), k * k * cin, k * k * cout)
        _bn(rng, W, scope, cout)
    c33 = num_chan_bn + 1
    W[ENC + "to_bn/weights"] = _xavier(rng, (5, 5, 128, c33), 25 * 128, 25 * c33)
    _bn(rng, W, ENC + "to_bn", c33)
    # decoder; transposed-conv filters are [k, k, out, in]
    W[DEC + "from_bn/weights"] = _xavier(rng, (3, 3, 128, num_chan_bn), 9 * 128, 9 * num_chan_bn)
    _bn(rng, W, DEC + "from_bn", 128)
    for scope, k, cin, cout in dec_conv_scopes(B):
        W[scope + "/weights"] = _xavier(rng, (k, k, cin, cout), k * k * cin, k * k * cout)
        _bn(rng, W, scope, cout)
    W[DEC + "h12/weights"] = _xavier(rng, (5, 5, 64, 128), 25 * 64, 25 * 128)
    _bn(rng, W, DEC + "h12", 64)
    W[DEC + "h13/weights"] = _xavier(rng, (5, 5, 3, 64), 25 * 3, 25 * 64)
    _bn(rng, W, DEC + "h13", 3)
    if residual_gamma is not None:
        for k_ in list(W):
            if k_.endswith("conv2/BatchNorm/gamma"):
                W[k_] = (W[k_] * residual_gamma).astype(np.float32)
    # probability classifier, conv3d filters DHWIO (src/probclass_imgcomp.py:227-261)
    L = num_centers
    for name, cin, cout in (("conv3d_conv0_mask", 1, pc_k), ("res1/conv3d_conv1_mask", pc_k, pc_k),
                            ("res1/conv3d_conv2_mask", pc_k, pc_k), ("conv3d_conv2_mask", pc_k, L)):
        W[PC + name + "/weights"] = _xavier(rng, (2, 3, 3, cin, cout), 18 * cin, 18 * cout)
        W[PC + name + "/biases"] = (0.05 * rng.standard_normal(cout)).astype(np.float32)
    # SI-Net: identity initialiser (src/siNet.py:13-20) plus small noise so that every
    # (dilated) tap participates in the parity check
    cin = 6
    for i, _rate in enumerate(SINET_RATES):
        w = np.zeros((3, 3, cin, 32), dtype=np.float32)
        for c in range(cin):
            w[1, 1, c, c] = 1.0
        w += 0.15 * _xavier(rng, w.shape, 9 * cin, 9 * 32)
        W[SIN + "g_conv%d/weights" % (i + 1)] = w
        W[SIN + "g_conv%d/biases" % (i + 1)] = (0.02 * rng.standard_normal(32)).astype(np.float32)
        cin = 32
    W[SIN + "g_conv_last/weights"] = _xavier(rng, (1, 1, 32, 3), 32, 3)
    W[SIN + "g_conv_last/biases"] = (0.02 * rng.standard_normal(3)).astype(np.float32)
    return W


_NAMES = {}


def variable_names(B=5):
    """TF variable names of the inference graph, in sorted order."""
    if B not in _NAMES:
        _NAMES[B] = sorted(make_weights(0, B=B))
    return list(_NAMES[B])


def set_bn_stats(W, scope, mean, var):
    W[scope + "/BatchNorm/moving_mean"] = np.asarray(mean, dtype=np.float32)
    W[scope + "/BatchNorm/moving_variance"] = np.maximum(np.asarray(var, dtype=np.float32), 1e-6)


def save_weights(path, W):
    np.savez(path, **{k.replace("/", "|"): v for k, v in W.items()})


def load_weights(path):
    with np.load(path) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}


# --------------------------------------------------------------------------------------
# synthetic stereo pairs
# --------------------------------------------------------------------------------------
def _smooth_field(rng, H, W, sigma):
    from scipy.ndimage import gaussian_filter
    f = gaussian_filter(rng.uniform(0, 1, size=(H, W)), sigma=sigma, mode="wrap")
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
    return f


def make_pair(seed, H=320, W=1224, sigma=None, disparity=None, noise=None):
    """One (x, y) pair, each (3,H,W) float32 holding uint8 values.  x is a smooth random
    field; y is x translated horizontally by ``disparity`` px plus Gaussian noise."""
    rng = np.random.default_rng(seed)
    sigma = float(rng.choice([2.0, 3.0, 6.0])) if sigma is None else sigma
    disparity = int(rng.integers(0, 65)) if disparity is None else int(disparity)
    noise = float(rng.uniform(2.0, 4.0)) if noise is None else noise
    Wp = W + disparity
    base = np.stack([_smooth_field(rng, H, Wp, sigma) for _ in range(3)], 0)
    # mix in a coarser component so there is large-scale structure too
    coarse = np.stack([_smooth_field(rng, H, Wp, 4 * sigma) for _ in range(3)], 0)
    img = 255.0 * (0.6 * base + 0.4 * coarse)
    x = img[:, :, disparity:disparity + W]
    y = img[:, :, 0:W] + noise * rng.standard_normal((3, H, W))
    x = np.clip(np.floor(x + 0.5), 0, 255).astype(np.float32)
    y = np.clip(np.floor(y + 0.5), 0, 255).astype(np.float32)
    return x, y


def make_batch(n, H=320, W=1224, seed=1000):
    xs, ys = zip(*(make_pair(seed + i, H, W) for i in range(n)))
    return np.stack(xs), np.stack(ys)
