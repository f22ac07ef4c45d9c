"""Generates tests/golden/*.npz from the parts of the reference that run in this container.

Run here (needs /root/reference):  python tests/golden/make_golden.py
  * mask_golden.npz   - AE.create_gaussian_masks (src/AE.py:193-220), extracted with ``ast``
                        and executed with a stub ``self`` (pure numpy).
  * msssim_golden.npz - ms_ssim_np_imgcomp.MultiScaleSSIM / utils.msssim_x_vs_rec call forms
                        (src/ms_ssim_np_imgcomp.py, src/utils.py:94-99) with ``tensorflow``
                        stubbed in sys.modules (the module never uses it).
Nothing from the reference is copied into the repo; only its numeric outputs are stored.
"""
import ast
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_mask_fn():
    src = open(os.path.join(REF, "AE.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "create_gaussian_masks":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, "AE.create_gaussian_masks", "exec"), ns)
            return ns["create_gaussian_masks"]
    raise RuntimeError("create_gaussian_masks not found")


def make_masks():
    fn = reference_mask_fn()
    out = {}
    for (H, W) in ((80, 144), (120, 96)):
        stub = types.SimpleNamespace(_y_patch_h=20, _y_patch_w=24, _input_dim_h=H, _input_dim_w=W)
        out["full_%dx%d" % (H, W)] = fn(stub)  # (1, h, w, P) float32
    rng = np.random.default_rng(7)
    for (H, W) in ((320, 1224), (320, 960)):
        stub = types.SimpleNamespace(_y_patch_h=20, _y_patch_w=24, _input_dim_h=H, _input_dim_w=W)
        m = fn(stub)[0]
        h, w, P = m.shape
        n = 4096
        ii, jj, pp = rng.integers(0, h, n), rng.integers(0, w, n), rng.integers(0, P, n)
        out["idx_%dx%d" % (H, W)] = np.stack([pp, ii, jj], 1).astype(np.int32)
        out["val_%dx%d" % (H, W)] = m[ii, jj, pp]
        out["argmax_%dx%d" % (H, W)] = m.reshape(h * w, P).argmax(0).astype(np.int64)
        out["shape_%dx%d" % (H, W)] = np.array(m.shape)
    np.savez_compressed(os.path.join(OUT, "mask_golden.npz"), **out)


def make_msssim():
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    sys.path.insert(0, REF)
    import ms_ssim_np_imgcomp as ref
    rng = np.random.default_rng(11)
    out = {}
    from scipy.ndimage import gaussian_filter
    for k, (H, W) in enumerate(((176, 208), (192, 256))):
        img = gaussian_filter(rng.uniform(0, 255, (H, W, 3)), (3, 3, 0))
        img = np.clip(255 * (img - img.min()) / (img.max() - img.min()), 0, 255).astype(np.uint8)
        rec8 = np.clip(np.round(img.astype(np.float32) + rng.normal(0, 4 + 4 * k, img.shape)), 0, 255).astype(np.uint8)
        rec = rec8.astype(np.float32)  # reconstructions are float arrays in utils.py
        out["img_%d" % k], out["rec_%d" % k] = img, rec8
        out["std_%d" % k] = np.float32(ref._calc_msssim_orig(img[None], rec[None]))
        # utils.msssim_x_vs_rec form: (H,W,3,1)
        out["utils_%d" % k] = np.float32(ref._calc_msssim_orig(img[..., None], rec[..., None]))
    np.savez_compressed(os.path.join(OUT, "msssim_golden.npz"), **out)


if __name__ == "__main__":
    make_masks()
    make_msssim()
    for f in ("mask_golden.npz", "msssim_golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))
