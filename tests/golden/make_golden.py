"""Generates tests/golden/*.npz from the parts of the reference that run in this container.

Run here (needs /root/reference):  python tests/golden/make_golden.py
  * mask_golden.npz   - AE.create_gaussian_masks (src/AE.py:193-220), extracted with ``ast``
                        and executed with a stub ``self`` (pure numpy).
  * msssim_golden.npz - ms_ssim_np_imgcomp.MultiScaleSSIM / utils.msssim_x_vs_rec call forms
                        (src/ms_ssim_np_imgcomp.py, src/utils.py:94-99) with ``tensorflow``
                        stubbed in sys.modules (the module never uses it).
  * model_pieces_golden.npz - the numpy-only pieces of the model code, extracted with ``ast`` and run with a stub
                        ``tf.name_scope``: the causal conv masks (src/probclass_imgcomp.py:150-176), the symbol
                        volume padding and its inverse (:268-292, :341-353), the block iteration order of the coder
                        helpers (:371-393), AE.normalize / denormalize / get_mean_var (src/AE.py:222-250), and the
                        host helpers readfiles (src/DataProvider.py:96-100), l1_x_vs_rec and save_test_imgs_fn
                        (src/utils.py:82-111).
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


def _extract(path, names):
    """Compile the named functions / methods of a reference file in isolation (no module import)."""
    import contextlib
    import functools
    import itertools
    tree = ast.parse(open(os.path.join(REF, path)).read())
    body = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in names]
    for n in body:
        n.decorator_list = []  # staticmethod / property decorators are meaningless outside the class
    tf_stub = types.SimpleNamespace(name_scope=lambda *_a, **_k: contextlib.nullcontext(), Variable=type(None))
    ns = {"np": np, "tf": tf_stub, "functools": functools, "itertools": itertools,
          "_Network3D": types.SimpleNamespace(_make_tf_conv3d_mask=lambda m: m)}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def make_model_pieces():
    pc = _extract("probclass_imgcomp.py", {
        "create_first_mask", "create_other_mask", "pad_for_probclass3d", "undo_pad_for_probclass3d", "get_np_pad_fn",
        "add_batch_dim", "remove_batch_dim", "_get_ndims", "iter_over_blocks", "num_blocks", "_iter_block_idices",
        "context_shape_from_context_size", "context_size_from_context_shape"})
    out = {}
    net = types.SimpleNamespace(config=types.SimpleNamespace(kernel_size=3), filter_shape=(2, 3, 3))
    out["first_mask"] = pc["create_first_mask"](net)    # (2,3,3,1,1)
    out["other_mask"] = pc["create_other_mask"](net)
    rng = np.random.default_rng(5)
    vol = rng.normal(size=(2, 3, 4, 5)).astype(np.float32)             # NCHW
    out["pad_in"] = vol
    out["pad_out_cs9"] = pc["pad_for_probclass3d"](vol, 9, 1.5)
    out["pad_out_chw_cs5"] = pc["pad_for_probclass3d"](vol[0], 5, -0.25)
    out["unpad_cs9"] = pc["undo_pad_for_probclass3d"](out["pad_out_cs9"][0], 9)  # CHW numpy branch
    out["context_shape_9"] = np.array(pc["context_shape_from_context_size"](9))
    syms = np.arange(4 * 5 * 6).reshape(4, 5, 6)
    blocks = list(pc["iter_over_blocks"](syms, (2, 3, 3)))
    out["block_first_elems"] = np.array([b[0, 0, 0] for b in blocks])
    out["block_count"] = np.array(pc["num_blocks"](syms.shape, (2, 3, 3)))
    ae = _extract("AE.py", {"normalize", "denormalize", "get_mean_var"})
    stub = types.SimpleNamespace(ae_config=types.SimpleNamespace(normalization="FIXED"), get_mean_var=ae["get_mean_var"])
    img = rng.integers(0, 256, size=(2, 3, 6, 7)).astype(np.float32)
    out["norm_in"] = img
    out["norm_out"] = ae["normalize"](stub, img)
    out["denorm_out"] = ae["denormalize"](stub, out["norm_out"])
    mean, var = ae["get_mean_var"]()
    out["mean"], out["var"] = mean, var
    # host-side helpers: pair-list reader (src/DataProvider.py:96-100), L1 metric and PNG writer (src/utils.py:82-111)
    import tempfile
    from PIL import Image
    dp = _extract("DataProvider.py", {"readfiles"})
    ut = _extract("utils.py", {"l1_x_vs_rec", "save_test_imgs_fn"})
    ut["os"], ut["Image"] = os, Image
    with tempfile.TemporaryDirectory() as tmp:
        lst = os.path.join(tmp, "pairs.txt")
        with open(lst, "w") as f:
            f.write("a/image_2/000000_10.png\na/image_3/000000_10.png\n  b/x.png  \nb/y.png\n")
        out["readfiles"] = np.array(dp["readfiles"](types.SimpleNamespace(root_data="/data/"), lst))
        rec = rng.uniform(-3, 260, size=(3, 5, 8)).astype(np.float32).clip(0, 255)
        ut["save_test_imgs_fn"](tmp + "/", "model", rec, 7, 0.0312345)
        name = os.listdir(os.path.join(tmp, "model"))
        out["png_name"] = np.array(name)
        out["png_in"] = rec
        out["png_pixels"] = np.asarray(Image.open(os.path.join(tmp, "model", name[0])))
    a = rng.integers(0, 256, size=(6, 7, 3)).astype(np.uint8)
    b = rng.uniform(0, 255, size=(6, 7, 3)).astype(np.float32)
    diff, l1 = ut["l1_x_vs_rec"](a, b)
    out["l1_a"], out["l1_b"], out["l1_diff"], out["l1_value"] = a, b, diff, np.float32(l1)
    np.savez_compressed(os.path.join(OUT, "model_pieces_golden.npz"), **out)


if __name__ == "__main__":
    make_masks()
    make_msssim()
    make_model_pieces()
    for f in ("mask_golden.npz", "msssim_golden.npz", "model_pieces_golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))
