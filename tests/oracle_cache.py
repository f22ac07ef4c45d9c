"""Cached oracle runs for the parity tests.  TEST INFRASTRUCTURE ONLY.

A *case* = (geometry, batch, seeds) -> BN-calibrated weights + the oracle's outputs for it: the fp32 restatement
(what the reference computes) and the float64 restatement of the encoder (used to adjudicate near-ties).
Cases are computed by oracle/dsin_oracle.py on first use and stored under tests/_cache/ (git-ignored; the
directory travels to the GPU box with the working tree, which saves the box from spending minutes of host
time per run).  The weights travel inside the case, so a cached case is self-consistent even though BN
calibration through torch-CPU is not bit-reproducible across machines.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from dsin_b200 import synth
from oracle import dsin_oracle as O
from oracle import ms_ssim_oracle as M

CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")
VERSION = 4


def _msssim_pair(x_chw, rec_chw):
    xi = np.transpose(x_chw, (1, 2, 0)).astype(np.uint8)
    ri = np.transpose(np.clip(rec_chw, 0, 255), (1, 2, 0))
    return float(M.msssim_standard(xi, ri)), float(M.msssim_reference_call(xi, ri))


def margins64(z64, centers):
    """|distance to the nearest centre - distance to the second nearest| per symbol, float64."""
    c = torch.as_tensor(np.asarray(centers), dtype=torch.float64)
    d = (z64.unsqueeze(-1) - c).abs()
    ds, _ = torch.sort(d, dim=-1)
    return ds[..., 1] - ds[..., 0]


def _bn_keys(Wt):
    return [k for k in Wt if "/BatchNorm/moving_" in k]


def compute_case(H, W, B, seed, wseed, keep_images=None, calib="std", light=False):
    Wt = synth.make_weights(wseed)
    x, y = synth.make_batch(B, H, W, seed=seed)
    if calib == "self":  # what __graft_entry__.smoke() does: calibrate on the case's own images
        O.calibrate_bn(Wt, np.concatenate([x, y]))
    else:
        xc, yc = synth.make_batch(2, 80, 144, seed=4242 + wseed)
        O.calibrate_bn(Wt, np.concatenate([xc, yc]))
    ref = O.reconstruct(x, y, Wt)
    out = {"x": x.astype(np.uint8), "y": y.astype(np.uint8)}
    keep = range(B) if keep_images is None else keep_images
    for k in ("y_dec", "y_syn", "x_dec", "x_with_si"):
        out["ref_" + k] = getattr(ref, k).numpy()[list(keep)]
    out["keep"] = np.array(list(keep), np.int64)
    out["ref_bpp"] = np.float64(float(ref.bpp))
    out["ref_symbols"] = ref.symbols.numpy().astype(np.int8)
    out["ref_row"], out["ref_col"] = ref.row.numpy(), ref.col.numpy()
    out["ref_best"] = ref.best.numpy()
    out["ref_bits_per_image"] = ref.bits_per_image.double().numpy()
    ms = [_msssim_pair(x[n], ref.x_with_si[n].numpy()) for n in range(B)]
    out["ref_msssim_std"] = np.array([m[0] for m in ms])
    out["ref_msssim_call"] = np.array([m[1] for m in ms])
    with torch.no_grad():
        e32x = O.encode(torch.as_tensor(x, dtype=torch.float32), Wt)
        e32y = O.encode(torch.as_tensor(y, dtype=torch.float32), Wt)
        e64x = O.encode(torch.as_tensor(x, dtype=torch.float64), Wt)
        e64y = O.encode(torch.as_tensor(y, dtype=torch.float64), Wt)
    c = Wt[O.ENC + "centers"]
    for tag, e32, e64 in (("x", e32x, e64x), ("y", e32y, e64y)):
        out["sym32_" + tag] = e32.symbols.numpy().astype(np.int8)
        out["sym64_" + tag] = e64.symbols.numpy().astype(np.int8)
        out["margin64_" + tag] = margins64(e64.z, c).numpy().astype(np.float32)
        if not light:  # the big per-symbol tensors only where a tool needs them
            out["z32_" + tag] = e32.z.numpy()
            out["z64_" + tag] = e64.z.numpy()
            out["qbar32_" + tag] = e32.qbar.numpy()
    # only the calibrated BN statistics travel: the rest of the weights is synth.make_weights(wseed) (pure numpy)
    for k in _bn_keys(Wt):
        out["W|" + k.replace("/", "|")] = Wt[k]
    out["wseed"] = np.int64(wseed)
    return out


def case(name, H, W, B, seed, wseed=0, keep_images=None, calib="std", light=False):
    """-> (weights dict, dict of numpy arrays).  Cached in tests/_cache/<name>.npz."""
    path = os.path.join(CACHE, "%s_v%d_%dx%d_b%d_s%d_w%d.npz" % (name, VERSION, H, W, B, seed, wseed))
    if not os.path.exists(path):
        os.makedirs(CACHE, exist_ok=True)
        d = compute_case(H, W, B, seed, wseed, keep_images, calib, light)
        np.savez(path + ".tmp.npz", **d)
        os.replace(path + ".tmp.npz", path)
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    Wt = synth.make_weights(int(d["wseed"]))
    Wt.update({k[2:].replace("|", "/"): v for k, v in d.items() if k.startswith("W|")})
    return Wt, {k: v for k, v in d.items() if not k.startswith("W|")}


# the cases the GPU tests and tools/precision_probe.py use
CASES = {
    "smoke": dict(H=80, W=144, B=1, seed=77, wseed=3, calib="self"),
    "small": dict(H=80, W=144, B=2, seed=300, wseed=0),
    "full1": dict(H=320, W=1224, B=1, seed=1000, wseed=0),
    "cfg4": dict(H=320, W=960, B=2, seed=4000, wseed=0),
    "full8": dict(H=320, W=1224, B=8, seed=2000, wseed=0, keep_images=[0, 5], light=True),
}


def get(name):
    return case(name, **CASES[name])


if __name__ == "__main__":
    import sys
    import time
    torch.set_num_threads(os.cpu_count() or 1)
    for nm in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        get(nm)
        print("case %s ready (%.1f s)" % (nm, time.time() - t0), flush=True)
