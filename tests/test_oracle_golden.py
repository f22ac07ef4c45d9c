"""Pins the oracle against outputs of the reference itself (tests/golden/*.npz, produced
by tests/golden/make_golden.py from /root/reference) and against closed-form invariants
implied by the reference code."""
import os

import numpy as np
import pytest
import torch

from oracle import dsin_oracle as O
from oracle import ms_ssim_oracle as M


def test_mask_matches_reference_full_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_golden.npz"))
    for (H, W) in ((80, 144), (120, 96)):
        ref = g["full_%dx%d" % (H, W)][0]  # (h, w, P)
        mine = O.gaussian_masks(H, W, 20, 24)  # (P, h, w)
        assert mine.dtype == np.float32
        assert np.array_equal(np.transpose(mine, (1, 2, 0)), ref)


@pytest.mark.parametrize("hw", [(320, 1224), (320, 960)])
def test_mask_matches_reference_sampled_full_size(golden_dir, hw):
    H, W = hw
    g = np.load(os.path.join(golden_dir, "mask_golden.npz"))
    mine = O.gaussian_masks(H, W, 20, 24)
    assert tuple(g["shape_%dx%d" % hw]) == (mine.shape[1], mine.shape[2], mine.shape[0])
    idx = g["idx_%dx%d" % hw]
    assert np.array_equal(mine[idx[:, 0], idx[:, 1], idx[:, 2]], g["val_%dx%d" % hw])
    am = mine.reshape(mine.shape[0], -1).argmax(1)
    assert np.array_equal(am, g["argmax_%dx%d" % hw])
    # the prior peaks at (top+1, left+1)
    w = mine.shape[2]
    assert (am[0] // w, am[0] % w) == (1, 1)
    assert (am[1] // w, am[1] % w) == (1, 25)


def test_msssim_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "msssim_golden.npz"))
    for k in (0, 1):
        img, rec = g["img_%d" % k], g["rec_%d" % k].astype(np.float32)
        assert M.msssim_standard(img, rec) == pytest.approx(float(g["std_%d" % k]), abs=1e-7)
        assert M.msssim_reference_call(img, rec) == pytest.approx(float(g["utils_%d" % k]), abs=1e-7)


def test_same_padding_rules():
    assert O.same_pads(320, 5, 2) == (1, 2)
    assert O.same_pads(320, 3, 2) == (0, 1)
    assert O.same_pads(306, 3, 1) == (1, 1)
    assert O.same_pads(1224, 3, 1, 128) == (128, 128)


def test_transposed_conv_is_adjoint_of_same_conv():
    """conv2d_transpose(SAME, s=2) == autograd adjoint of the SAME s=2 conv."""
    rng = np.random.default_rng(0)
    for k in (3, 5):
        w = rng.standard_normal((k, k, 4, 6)).astype(np.float64)  # [k,k,out(=4),in(=6)]
        q = torch.tensor(rng.standard_normal((1, 6, 5, 7)))
        x = torch.zeros(1, 4, 10, 14, dtype=torch.float64, requires_grad=True)
        y = O.conv2d_same(x, w, stride=2)  # HWIO with I=4 (x channels), O=6
        (g,) = torch.autograd.grad(y, x, grad_outputs=q)
        mine = O.conv2d_transpose_same_s2(q, w)
        assert torch.allclose(mine, g, atol=1e-12)


def test_quantizer_picks_nearest_centre_first_index():
    c = np.array([0.5, -1.0, 0.5, 2.0], dtype=np.float32)
    z = torch.tensor([[[[0.5, -0.9, 1.9, 0.6]]]])
    qbar, qsoft, qhard, sym = O.quantize(z, c)
    assert sym.dtype == torch.int64
    assert sym.flatten().tolist() == [0, 1, 3, 0]  # duplicate centre -> first index
    assert torch.allclose(qhard.flatten(), torch.tensor([0.5, -1.0, 2.0, 0.5]))
    assert torch.allclose(qbar, qhard, atol=1e-6)


def test_heatmap_zero_gives_constant_symbol():
    W = {O.ENC + "centers": np.array([0.7, -0.2, 1.5, -1.1, 0.1, 2.0], dtype=np.float32)}
    z33 = torch.randn(1, 33, 4, 5)
    z33[:, 0] = -100.0  # sigmoid -> 0 -> heatmap3D == 0
    hm = O.heatmap3d(z33)
    assert float(hm.abs().max()) == 0.0
    _, _, _, sym = O.quantize(hm * z33[:, 1:], W[O.ENC + "centers"])
    assert (sym == 4).all()  # argmin |c_j| = 0.1


def test_patch_tiling_roundtrip():
    img = torch.arange(40 * 48 * 3, dtype=torch.float32).reshape(40, 48, 3)
    p = O.extract_patches(img, 20, 24)
    assert p.shape == (4, 20, 24, 3)
    assert torch.equal(p[1], img[0:20, 24:48])
    assert torch.equal(O.fold_patches(p, 40, 48), img)


def test_sifinder_identity_pair_finds_own_origin():
    """y == x: Pearson == 1 at the patch origin; the mask peak sits at origin+(1,1), so
    the argmax is at the origin or within one pixel of it."""
    from dsin_b200 import synth
    x, _ = synth.make_pair(3, 80, 144, sigma=2.0, disparity=0, noise=0.0)
    xt = torch.tensor(x)[None]
    y_syn, row, col, best = O.si_full_img(xt, xt, xt)
    P = row.shape[1]
    for p in range(P):
        pr, pc = (p // 6) * 20, (p % 6) * 24
        assert abs(int(row[0, p]) - pr) <= 1 and abs(int(col[0, p]) - pc) <= 1
    assert float(best.min()) > 0.99


def test_probclass_is_causal():
    from dsin_b200 import synth
    W = synth.make_weights(1)
    rng = np.random.default_rng(0)
    c = W[O.ENC + "centers"]
    sym = torch.tensor(rng.integers(0, 6, (1, 8, 6, 7)))
    q = torch.tensor(c)[sym]
    b0 = O.probclass_bitcost(q, sym, W)
    # change one voxel: only voxels at or after it in (C,H,W) raster order may change
    sym2 = sym.clone()
    sym2[0, 3, 2, 4] = (sym2[0, 3, 2, 4] + 1) % 6
    q2 = torch.tensor(c)[sym2]
    b1 = O.probclass_bitcost(q2, sym2, W)
    changed = (b0 != b1)[0]
    order = torch.arange(8 * 6 * 7).reshape(8, 6, 7)
    assert changed.any()
    assert int(order[changed].min()) >= int(order[3, 2, 4])


def test_sinet_identity_init_passes_channels_through():
    """src/siNet.py:13-20: identity-initialised 3x3 layers copy their input channels."""
    W = {}
    cin = 6
    for i in range(9):
        w = np.zeros((3, 3, cin, 32), dtype=np.float32)
        for c in range(cin):
            w[1, 1, c, c] = 1
        W[O.SIN + "g_conv%d/weights" % (i + 1)] = w
        W[O.SIN + "g_conv%d/biases" % (i + 1)] = np.zeros(32, dtype=np.float32)
        cin = 32
    last = np.zeros((1, 1, 32, 3), dtype=np.float32)
    for c in range(3):
        last[0, 0, c, c] = 1
    W[O.SIN + "g_conv_last/weights"] = last
    W[O.SIN + "g_conv_last/biases"] = np.zeros(3, dtype=np.float32)
    x = torch.rand(1, 6, 16, 20) + 0.1  # positive: lrelu is the identity
    out = O.si_net(x, W)
    assert torch.allclose(out, x[:, :3], atol=1e-6)


def test_crop_and_resize_coordinates():
    """sampling coordinates for row 100 of a 320-row image."""
    y = np.tile(np.arange(320, dtype=np.float32)[:, None, None], (1, 1224, 3))
    out = O.crop_and_resize_patches(y, [100], [0], 20, 24)
    assert out[0, 0, 0, 0] == pytest.approx(99.6875, abs=1e-4)
    assert out[0, 1, 0, 0] == pytest.approx(100.7368, abs=1e-3)
    assert out[0, 2, 0, 0] == pytest.approx(101.7862, abs=1e-3)


def test_model_pieces_match_reference(golden_dir):
    """Numpy-only pieces of the reference's model code, executed from the reference itself by
    tests/golden/make_golden.py: causal masks, symbol-volume padding, normalisation, coder block order."""
    import torch.nn.functional as F
    from dsin_b200 import probclass_imgcomp as product_pc
    from oracle import pc_codec as P
    g = np.load(os.path.join(golden_dir, "model_pieces_golden.npz"))
    first, other = O.pc_masks(3)
    assert np.array_equal(first, g["first_mask"][..., 0, 0]) and np.array_equal(other, g["other_mask"][..., 0, 0])
    pf, po = product_pc.create_masks(3)
    assert np.array_equal(pf, first) and np.array_equal(po, other)
    # the entropy coder's live-tap lists are exactly the non-zero mask entries in raster (kd, kh, kw) order
    live = lambda m: [tuple(int(v) for v in idx) for idx in np.argwhere(m != 0)]  # noqa: E731
    assert P.TAPS_FIRST == live(first) and P.TAPS_OTHER == live(other)
    # padding of the symbol volume: front of the depth axis and both sides of H, W; nothing behind in depth
    x = torch.from_numpy(g["pad_in"])
    assert np.array_equal(F.pad(x, (4, 4, 4, 4, 4, 0), value=1.5).numpy(), g["pad_out_cs9"])
    assert np.array_equal(F.pad(x[0], (2, 2, 2, 2, 2, 0), value=-0.25).numpy(), g["pad_out_chw_cs5"])
    assert np.array_equal(g["unpad_cs9"], g["pad_in"][0]) and list(g["context_shape_9"]) == [5, 9, 9]
    # normalisation constants and arithmetic (float32, sqrt(var + 1e-10) evaluated in float32)
    assert np.array_equal(O.KITTI_MEAN, g["mean"].reshape(3)) and np.array_equal(O.KITTI_VAR, g["var"].reshape(3))
    from dsin_b200.AE import AE
    pm, pv = AE.get_mean_var()
    assert np.array_equal(pm, g["mean"]) and np.array_equal(pv, g["var"])
    n = O.normalize(torch.from_numpy(g["norm_in"]))
    assert np.array_equal(n.numpy(), g["norm_out"])
    assert np.array_equal(O.denormalize(torch.from_numpy(g["norm_out"])).numpy(), g["denorm_out"])
    # the coder helpers walk the symbol volume with W fastest, then H, then C (src/probclass_imgcomp.py:383-393)
    want = [c * 30 + h * 6 + w for c in range(3) for h in range(3) for w in range(4)]
    assert g["block_first_elems"].tolist() == want and int(g["block_count"]) == 36


def test_host_helpers_match_reference(golden_dir, tmp_path):
    """Pair-list reader, PNG writer (name + uint8 truncation) and L1 metric against the reference's own functions
    (src/DataProvider.py:96-100, src/utils.py:82-111), executed by tests/golden/make_golden.py."""
    import types
    from PIL import Image
    from dsin_b200 import utils
    from dsin_b200.DataProvider import Dataset
    g = np.load(os.path.join(golden_dir, "model_pieces_golden.npz"))
    lst = tmp_path / "pairs.txt"
    lst.write_text("a/image_2/000000_10.png\na/image_3/000000_10.png\n  b/x.png  \nb/y.png\n")
    got = Dataset.readfiles(types.SimpleNamespace(root_data="/data/"), str(lst))
    assert got == g["readfiles"].tolist()
    utils.save_test_imgs_fn(str(tmp_path) + "/", "model", g["png_in"], 7, 0.0312345)
    assert os.listdir(tmp_path / "model") == g["png_name"].tolist()
    assert np.array_equal(np.asarray(Image.open(tmp_path / "model" / str(g["png_name"][0]))), g["png_pixels"])
    diff, l1 = utils.l1_x_vs_rec(g["l1_a"], g["l1_b"])
    assert np.array_equal(diff, g["l1_diff"]) and np.float32(l1) == g["l1_value"]


# ----------------------------------------------------------------------------- reference code under a numpy tf stand-in
def test_oracle_matches_reference_elementwise_functions(golden_dir):
    """tests/golden/tf_pieces_golden.npz holds outputs of the reference's OWN functions (ast-extracted, executed under a
    numpy stand-in for the elementwise / shape tf ops they call; tests/golden/make_golden.py): SI-Finder normalisation and
    colour transform (src/siFinder.py:56-73,138-154), heatmap (src/autoencoder_imgcomp.py:173-201), quantiser
    (src/quantizer_imgcomp.py:43-100), bpp (src/bits_imgcomp.py:4-20).  The oracle must reproduce them."""
    g = np.load(os.path.join(golden_dir, "tf_pieces_golden.npz"))
    img = torch.tensor(g["sif_in"])
    norm = O.sif_normalize_nhwc(img)
    assert torch.equal(norm, torch.tensor(g["sif_norm"]))
    assert torch.equal(O.rgb_transform(norm), torch.tensor(g["sif_rgb"]))
    z33 = torch.tensor(g["z33"])
    hm = O.heatmap3d(z33)
    assert float((hm - torch.tensor(g["heatmap3d"])).abs().max()) <= 4e-6  # sigmoid: torch vs 1/(1+exp(-x)) in fp32
    assert float((hm * z33[:, 1:] - torch.tensor(g["z_masked"])).abs().max()) <= 2e-5
    qbar, qsoft, qhard, sym = O.quantize(torch.tensor(g["q_in"]), g["q_centers"])
    assert torch.equal(sym, torch.tensor(g["q_symbols"]))   # incl. the exactly-on-a-centre and the midway cases
    assert torch.equal(qhard, torch.tensor(g["q_hard"]))
    assert float((qsoft - torch.tensor(g["q_soft"])).abs().max()) <= 5e-7
    n_pix = int(np.prod(g["bc_input_shape"])) // 3
    assert float(O.bitcost_to_bpp(torch.tensor(g["bc"]), n_pix)) == pytest.approx(float(g["bpp"]), rel=2e-7)


def test_loss_arithmetic_matches_reference(golden_dir):
    """siNet_validate's loss arithmetic (src/AE.py:76-99): the oracle's restatement AND the product's host-side float32
    arithmetic (dsin_b200/Distortions_imgcomp.py, fed with float64 sums like the ones csrc/loss.cu produces) against the
    reference's own Distortions class and get_loss, executed on the numpy stand-in for TF (make_golden.make_loss_pieces)."""
    import types

    from dsin_b200 import Distortions_imgcomp as D
    g = np.load(os.path.join(golden_dir, "loss_pieces_golden.npz"))
    x, xo, bc, hm = (g[k] for k in ("x", "x_out", "bc", "heatmap"))
    tx, txo, tbc, thm = (torch.tensor(a) for a in (x, xo, bc, hm))
    n = x.shape[0]
    img_elems = x[0].size
    ae = types.SimpleNamespace(encoder_regularization_loss=lambda: np.float32(g["reg_enc"]),
                               decoder_regularization_loss=lambda: np.float32(g["reg_dec"]))
    pc = types.SimpleNamespace(regularization_loss=lambda: None)
    H_real = np.float32(bc.astype(np.float64).sum() / bc.size)
    H_mask = np.float32((bc.astype(np.float32) * hm).astype(np.float64).sum() / bc.size)
    assert abs(float(H_real) - float(g["H_real"])) <= 2e-7 and abs(float(H_mask) - float(g["H_mask"])) <= 2e-7
    for kind in ("mae", "mse", "psnr"):
        d_or = O.distortion_to_minimize(tx, txo, kind, 100.0)
        assert abs(float(d_or) - float(g["d_" + kind])) <= 2e-6 * abs(float(g["d_" + kind]))
        diff = (xo.astype(np.float32) - x).astype(np.float64)
        sums = (np.abs(diff) if kind == "mae" else diff * diff).reshape(n, -1).sum(1)
        cfg = types.SimpleNamespace(distortion_to_minimize=kind, K_psnr=100, beta=500, H_target=0.04)
        assert D.squared_distortion(cfg) == (kind != "mae")
        d_pr = D.distortion_to_minimize(cfg, sums, img_elems)
        assert abs(float(d_pr) - float(g["d_" + kind])) <= 2e-6 * abs(float(g["d_" + kind]))
        for h_target in (0.04, 2.5):
            want = float(g["total_%s_%g" % (kind, h_target)])
            tot, hr, hk, _pl = O.get_loss(np.float32(0.3) * d_or, tbc, thm, 500.0, h_target, float(g["reg_enc"]),
                                          float(g["reg_dec"]))
            assert abs(float(tot) - want) <= 2e-6 * abs(want)
            assert abs(float(hr) - float(g["H_real"])) <= 2e-7 and abs(float(hk) - float(g["H_mask"])) <= 2e-7
            cfg.H_target = h_target
            tot_p, hr_p, pc_comps, ae_comps = D.get_loss(cfg, ae, pc, np.float32(0.3) * d_pr, H_real, H_mask)
            assert abs(float(tot_p) - want) <= 2e-6 * abs(want)
            assert dict(pc_comps)["reg"] == 0 and abs(float(dict(ae_comps)["reg_enc_dec"]) - 0.5) < 1e-7
    with pytest.raises(NotImplementedError):
        D.distortion_to_minimize(types.SimpleNamespace(distortion_to_minimize="ms_ssim"), [1.0], 10)


def test_regularization_scope_rule():
    """tf.losses.get_regularization_loss(scope) filters by re.match on the op name (a prefix match): with the variable
    names of the graph src/AE.py builds nothing matches 'autoencoder/encoder' / 'autoencoder/decoder' and the term is
    0.0; weights that DO live under those scopes are summed as factor * sum(w^2) / 2 (+ the centres' term)."""
    from dsin_b200 import Distortions_imgcomp as D
    from dsin_b200 import synth
    W = synth.make_weights(0)
    for fn in (D.regularization_loss, O.regularization_loss):
        assert float(fn(W, "autoencoder/encoder", 0.005, 0.1)) == 0.0
        assert float(fn(W, "autoencoder/decoder", 0.005, 0.0)) == 0.0
        renamed = {k[len("encoder/encoder_body/encoder_body/"):]: v for k, v in W.items() if k.startswith(O.ENC)}
        want = 0.0
        for k, v in renamed.items():
            v = np.asarray(v, np.float64)
            if k.endswith("/weights"):
                want += 0.005 * 0.5 * float((v * v).sum())
            elif k.endswith("/centers"):
                want += 0.1 * 0.5 * float((v * v).sum())
        got = float(fn(renamed, "autoencoder/encoder", 0.005, 0.1))
        assert want > 0 and abs(got - want) <= 1e-6 * want
        assert float(fn(renamed, "autoencoder/decoder", 0.005, 0.0)) == 0.0
