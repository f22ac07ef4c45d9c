"""GPU parity tests: every kernel of libdsin_b200 (through the C ABI) against the CPU oracle on
identical seeded inputs/weights.  Integer outputs (symbols, SI-Finder row/col) must be equal
except at near-ties adjudicated by the float64 oracle; floats within the stated tolerances
(north_star: |d bpp| <= 1e-5, |d MS-SSIM| <= 1e-4)."""
import os

import numpy as np
import pytest
import torch

from dsin_b200 import synth
from oracle import dsin_oracle as O
from oracle import ms_ssim_oracle as M  # noqa: F401

from parity_utils import calibrated_weights, make_ae, symbol_report

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


# ----------------------------------------------------------------------------- K1/K2/K8 conv
@pytest.mark.parametrize("case", [
    dict(k=3, cin=128, cout=128, stride=1, dil=1, tr=False, h=20, w=38),
    dict(k=5, cin=3, cout=64, stride=2, dil=1, tr=False, h=40, w=72),
    dict(k=5, cin=64, cout=128, stride=2, dil=1, tr=False, h=20, w=36),
    dict(k=5, cin=128, cout=33, stride=2, dil=1, tr=False, h=20, w=18),
    dict(k=3, cin=32, cout=128, stride=2, dil=1, tr=True, h=10, w=19),
    dict(k=5, cin=128, cout=64, stride=2, dil=1, tr=True, h=10, w=18),
    dict(k=5, cin=64, cout=3, stride=2, dil=1, tr=True, h=20, w=36),
    dict(k=3, cin=6, cout=32, stride=1, dil=1, tr=False, h=40, w=48),
    dict(k=3, cin=32, cout=32, stride=1, dil=16, tr=False, h=40, w=48),
    dict(k=3, cin=32, cout=32, stride=1, dil=128, tr=False, h=40, w=48),
    dict(k=1, cin=32, cout=3, stride=1, dil=1, tr=False, h=24, w=40),
])
def test_conv2d_matches_oracle(case):
    from dsin_b200 import ops
    rng = np.random.default_rng(1)
    k, cin, cout = case["k"], case["cin"], case["cout"]
    x = rng.standard_normal((2, cin, case["h"], case["w"])).astype(np.float32)
    if case["tr"]:
        w_ref = (rng.standard_normal((k, k, cout, cin)) / np.sqrt(k * k * cin)).astype(np.float32)
        w_pack = np.transpose(w_ref, (0, 1, 3, 2))
        ref = O.conv2d_transpose_same_s2(torch.tensor(x), w_ref)
    else:
        w_ref = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
        w_pack = w_ref
        ref = O.conv2d_same(torch.tensor(x), w_ref, stride=case["stride"], dilation=case["dil"])
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal(tuple(ref.shape)).astype(np.float32)
    ref = torch.relu(ref * torch.tensor(scale).view(1, -1, 1, 1) + torch.tensor(shift).view(1, -1, 1, 1)) + torch.tensor(res)
    layer = ops.ConvLayer(w_pack, scale, shift, stride=case["stride"], dilation=case["dil"], transposed=case["tr"],
                          act=ops.ACT_RELU)
    got = ops.conv2d(_nhwc(_dev(x)), layer, res1=_nhwc(_dev(res)))
    got = got.permute(0, 3, 1, 2).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def test_layout_and_normalise_bit_exact():
    from dsin_b200 import ops
    x, _ = synth.make_batch(2, 40, 48, seed=3)
    got = ops.nchw_to_nhwc(_dev(x), normalize=True).cpu()
    ref = O.normalize(torch.tensor(x)).permute(0, 2, 3, 1)
    assert torch.equal(got, ref)
    back = ops.nhwc_to_nchw(ops.nchw_to_nhwc(_dev(x))).cpu()
    assert torch.equal(back, torch.tensor(x))


# ----------------------------------------------------------------------------- K3 quantiser
def test_quantizer_bit_exact_on_same_input():
    from dsin_b200 import ops
    rng = np.random.default_rng(2)
    z33 = (1.5 * rng.standard_normal((2, 33, 10, 19))).astype(np.float32)
    z33[0, 0, :2] = -100.0  # heatmap off -> constant symbol
    z33[0, 1:, 3, 3] = 0.0
    centers = np.array([0.55, -0.92, -1.84, -1.93, 1.25, 1.65], dtype=np.float32)
    hm = O.heatmap3d(torch.tensor(z33))
    qbar, _qs, _qh, sym = O.quantize(hm * torch.tensor(z33)[:, 1:], centers)
    q_nhwc, q_nchw, s = ops.heatmap_quantize(_nhwc(_dev(z33)), _dev(centers))
    assert s.dtype == torch.int64
    assert torch.equal(s.cpu(), sym)  # integer symbol indices: bit-exact
    assert float((q_nchw.cpu() - qbar).abs().max()) <= 4e-7
    assert torch.equal(q_nhwc.permute(0, 3, 1, 2).cpu(), q_nchw.cpu())


# ----------------------------------------------------------------------------- K4 probclass
def test_probclass_bits_match_oracle():
    W = calibrated_weights(0)
    ae = make_ae(80, 144, W)
    rng = np.random.default_rng(5)
    c = W[O.ENC + "centers"]
    sym = torch.tensor(rng.integers(0, 6, (3, 32, 10, 18)))
    q = torch.tensor(c)[sym]
    ref = O.probclass_bitcost(q, sym, W)
    bits = ae.pc_imgcomp.bitcost(q.cuda(), sym.cuda(), is_training=False, pad_value=float(c[0]))
    assert float((bits.cpu() - ref).abs().max()) < 1e-4  # fp32 accumulation order, 4 layers K=432
    sums = bits._dsin_sum.cpu()
    assert torch.allclose(sums, ref.double().reshape(3, -1).sum(1), rtol=1e-6)


# ----------------------------------------------------------------------------- K5-K7 SI-Finder
def _sif_case(H, W, seed, n=2):
    xs, ys = [], []
    for i in range(n):
        x, y = synth.make_pair(seed + i, H, W, sigma=2.0, disparity=8 + 5 * i, noise=3.0)
        xs.append(x)
        ys.append(y)
    return np.stack(xs), np.stack(ys)


@pytest.fixture(params=[0, 1], ids=["simt", "tcgen05"])
def sif_method(request):
    from dsin_b200 import siFinder as sf
    old = sf.METHOD
    sf.METHOD = request.param
    yield request.param
    sf.METHOD = old


@pytest.mark.parametrize("hw", [(80, 144), (120, 96)])
def test_sifinder_rowcol_and_gather_match_oracle(hw, sif_method):
    from dsin_b200.siFinder import match_images
    H, W = hw
    x, y = _sif_case(H, W, 10)
    # feed the images directly as "decoded" inputs: structured, high-correlation matches
    xt, yt = torch.tensor(x), torch.tensor(y)
    ref_syn, ref_row, ref_col, ref_best = O.si_full_img(xt, yt, yt)
    y_syn, q, r, row, col, best = match_images(_nhwc(_dev(x)), _nhwc(_dev(y)), _nhwc(_dev(y)), 20, 24, True)
    row, col = row.cpu(), col.cpu()
    mism = (row != ref_row) | (col != ref_col)
    # adjudicate mismatches with the float64 oracle: must be near-ties
    if int(mism.sum()):
        mask = O.gaussian_masks(H, W, 20, 24)
        for n, p in zip(*np.nonzero(mism.numpy())):
            xi = torch.tensor(x[n], dtype=torch.float64).permute(1, 2, 0)
            yi = torch.tensor(y[n], dtype=torch.float64).permute(1, 2, 0)
            q64 = O.rgb_transform(O.sif_normalize_nhwc(O.extract_patches(xi, 20, 24)))
            r64 = O.rgb_transform(O.sif_normalize_nhwc(yi))
            a = O.score_at(q64, r64, mask, p, int(row[n, p]), int(col[n, p]))
            b = O.score_at(q64, r64, mask, p, int(ref_row[n, p]), int(ref_col[n, p]))
            assert abs(a - b) < 2e-6, (n, p, a, b)
    assert int(mism.sum()) <= 1
    same = ~mism
    assert float((best.cpu() - ref_best)[same].abs().max()) < 2e-5
    got_syn = y_syn.permute(0, 3, 1, 2).cpu()
    for n in range(x.shape[0]):
        if not bool(mism[n].any()):
            assert torch.equal(got_syn[n], ref_syn[n])  # bilinear gather: bit-exact fp32


def _adjudicate_rowcol(x_dec, y_dec, row, col, ref_row, ref_col, tol=2e-6):
    """Every (row,col) that differs from the fp32 oracle must be a near-tie in float64."""
    mism = (row != ref_row) | (col != ref_col)
    N, _, H, W = x_dec.shape
    if int(mism.sum()):
        mask = O.gaussian_masks(H, W, 20, 24)
        for n, p in zip(*np.nonzero(mism.numpy())):
            xi = torch.as_tensor(x_dec[n], dtype=torch.float64).permute(1, 2, 0)
            yi = torch.as_tensor(y_dec[n], dtype=torch.float64).permute(1, 2, 0)
            q64 = O.rgb_transform(O.sif_normalize_nhwc(O.extract_patches(xi, 20, 24)))
            r64 = O.rgb_transform(O.sif_normalize_nhwc(yi))
            a = O.score_at(q64, r64, mask, p, int(row[n, p]), int(col[n, p]))
            b = O.score_at(q64, r64, mask, p, int(ref_row[n, p]), int(ref_col[n, p]))
            assert abs(a - b) < tol, (n, p, a, b)
    return mism


def test_sifinder_tc_equals_simt_full_size():
    """320x1224, 2 structured pairs: tcgen05 coarse + exact rescoring vs the fp32 SIMT scorer."""
    from dsin_b200 import siFinder as sf
    x, y = _sif_case(320, 1224, 40)
    args = (_nhwc(_dev(x)), _nhwc(_dev(y)), _nhwc(_dev(y)), 20, 24, True)
    old = sf.METHOD
    try:
        sf.METHOD = 0
        _, _, _, row0, col0, best0 = sf.match_images(*args)
        sf.METHOD = 1
        _, _, _, row1, col1, best1 = sf.match_images(*args)
    finally:
        sf.METHOD = old
    mism = _adjudicate_rowcol(x, y, row1.cpu(), col1.cpu(), row0.cpu(), col0.cpu(), tol=3e-6)
    print("tc vs simt: %d / %d positions differ (all float64 near-ties)" % (int(mism.sum()), mism.numel()))
    assert int(mism.sum()) <= 4
    same = ~mism
    assert float((best1.cpu() - best0.cpu())[same].abs().max()) < 2e-5


def test_sifinder_flat_rgb_patch_matches_oracle(sif_method):
    """A patch that is constant in RGB is NOT degenerate: after the per-channel normalisation and
    colour transform its 1440-vector still varies across channels (den_x > 0)."""
    from dsin_b200.siFinder import match_images
    x, y = _sif_case(80, 144, 20, n=1)
    x[0, :, 20:40, 24:48] = 255.0  # patch p = 1*6+1 = 7
    xt, yt = torch.tensor(x), torch.tensor(y)
    _, ref_row, ref_col, ref_best = O.si_full_img(xt, yt, yt)
    _, _, _, row, col, best = match_images(_nhwc(_dev(x)), _nhwc(_dev(y)), _nhwc(_dev(y)), 20, 24, True)
    mism = _adjudicate_rowcol(x, y, row.cpu(), col.cpu(), ref_row, ref_col, tol=1e-5)
    assert int(mism.sum()) <= 1
    assert bool(torch.isfinite(best).all())


# ----------------------------------------------------------------------------- stage-wise, full pipeline
def _stagewise(H, W, B, seed):
    """Each stage of AE.siNet_get_reconstructed is fed the ORACLE's input for that stage, so one
    near-tie flip cannot cascade through the (chaotic, random-init) later stages."""
    from dsin_b200.siFinder import match_images
    Wt = calibrated_weights(0)
    ae = make_ae(H, W, Wt)
    x, y = synth.make_batch(B, H, W, seed=seed)
    ref = O.reconstruct(x, y, Wt)
    rep = {}
    # (a) encoder + quantiser: integer symbols
    enc = ae.ae_imgcomp.encode(_dev(x))
    n_mism, bad, total = symbol_report(enc.symbols.cpu(), x, Wt)
    assert bad == 0, "symbol mismatches that are not float64 near-ties: %d" % bad
    assert n_mism <= max(2, total // 100000), (n_mism, total)
    rep["symbol_mismatch"] = (n_mism, total)
    # (b) probability model on the oracle's qbar/symbols
    enc_ref = O.encode(torch.tensor(x), Wt)
    bc = ae.pc_imgcomp.bitcost(enc_ref.qbar.cuda().contiguous(), enc_ref.symbols.cuda(), False,
                               pad_value=ae.pc_imgcomp.auto_pad_value(ae.ae_imgcomp))
    bpp_gpu = float(bc._dsin_sum.sum().item()) / (B * H * W)
    assert abs(bpp_gpu - float(ref.bpp)) <= 1e-5, (bpp_gpu, float(ref.bpp))
    rep["bpp"] = (bpp_gpu, float(ref.bpp))
    # (c) decoder on the oracle's qbar
    x_dec = ae.ae_imgcomp.decode(enc_ref.qbar.cuda().contiguous()).cpu()
    err = float((x_dec - ref.x_dec).abs().max())
    assert err < 2e-2, err  # 0..255 scale
    rep["x_dec_err"] = err
    # (d) SI-Finder on the oracle's x_dec / y_dec
    y_syn, _q, _r, row, col, best = match_images(_nhwc(ref.x_dec.cuda()), _nhwc(_dev(y)), _nhwc(ref.y_dec.cuda()),
                                                 20, 24, True)
    mism = _adjudicate_rowcol(ref.x_dec.numpy(), ref.y_dec.numpy(), row.cpu(), col.cpu(), ref.row, ref.col)
    assert int(mism.sum()) <= max(1, mism.numel() // 400)
    rep["rowcol_mismatch"] = (int(mism.sum()), mism.numel())
    got_syn = y_syn.permute(0, 3, 1, 2).cpu()
    for n in range(B):
        if not bool(mism[n].any()):
            assert torch.equal(got_syn[n], ref.y_syn[n])
    # (e) SI-Net on the oracle's x_dec / y_syn
    xsi = ae._siNet.fused(_nhwc(ref.x_dec.cuda()), _nhwc(ref.y_syn.cuda())).cpu()
    err = float((xsi - ref.x_with_si).abs().max())
    assert err < 2e-2, err
    rep["x_with_si_err"] = err
    return ae, x, y, ref, rep


def test_stagewise_small():
    _stagewise(80, 144, 2, 300)


def test_stagewise_full_size():
    """BASELINE config 1/2 geometry (320x1224), every stage fed the oracle's input for that stage.  The free-running
    public call is held to the north_star tolerances in tests/test_gpu_freerun.py."""
    ae, x, y, ref, rep = _stagewise(320, 1224, 1, 1000)
    print("full-size stage-wise report:", rep)
    # the shipped policy runs the decoders on fp16 operands (precision.py): dense sub-grey-level noise, no outliers
    enc_ref = O.encode(torch.tensor(x), calibrated_weights(0))
    x_dec1 = ae.ae_imgcomp.decode(enc_ref.qbar.cuda().contiguous(), terms=1).cpu()
    diff = (x_dec1 - ref.x_dec).abs()
    print("fp16-operand decoder vs oracle: max %.3f rms %.4f grey levels" % (float(diff.max()), float(diff.pow(2).mean().sqrt())))
    assert float(diff.max()) < 1.5 and float(diff.pow(2).mean().sqrt()) < 0.15


# ----------------------------------------------------------------------------- tcgen05 trunk conv
@pytest.mark.parametrize("terms,tol", [(3, 1e-5), (1, 3e-3)])  # 1e-5: fp32 accumulation over K=1152
@pytest.mark.parametrize("shape", [(2, 20, 36), (1, 80, 306), (3, 9, 17)])
def test_conv3x3_tc_matches_fp32_conv(terms, tol, shape):
    """tcgen05 3x3 128->128 conv (split-fp16) vs the oracle's fp32 conv on the same (split-rounded) input."""
    from dsin_b200 import ops
    rng = np.random.default_rng(7)
    n, hh, ww = shape
    x = rng.standard_normal((n, 128, hh, ww)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 128, 128)) / np.sqrt(9 * 128)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 128).astype(np.float32)
    shift = rng.standard_normal(128).astype(np.float32)
    r1 = rng.standard_normal((n, 128, hh, ww)).astype(np.float32)
    r2 = rng.standard_normal((n, 128, hh, ww)).astype(np.float32)
    layer = ops.ConvLayer(w, scale, shift, act=ops.ACT_RELU)
    tcl = ops.Conv3x3TC(layer)
    xs = ops.f32_to_split(_nhwc(_dev(x)))
    r1s = ops.f32_to_split(_nhwc(_dev(r1)))
    r2s = ops.f32_to_split(_nhwc(_dev(r2)))
    yh, yl = ops.conv3x3_tc(xs[0], xs[1], tcl, res1=r1s, res2=r2s, terms=terms)
    got = ops.split_to_f32(yh, yl).permute(0, 3, 1, 2).cpu().double()
    # reference in float64 from the values the kernel actually saw
    xq = ops.split_to_f32(*xs).permute(0, 3, 1, 2).cpu().double()
    r1q = ops.split_to_f32(*r1s).permute(0, 3, 1, 2).cpu().double()
    r2q = ops.split_to_f32(*r2s).permute(0, 3, 1, 2).cpu().double()
    ref = O.conv2d_same(xq, w.astype(np.float64))
    ref = torch.relu(ref * torch.tensor(scale).double().view(1, -1, 1, 1) + torch.tensor(shift).double().view(1, -1, 1, 1))
    ref = ref + r1q + r2q
    err = float((got - ref).abs().max())
    assert err < tol * max(1.0, float(ref.abs().max())), err


@pytest.mark.parametrize("terms,tol", [(3, 2e-5), (1, 4e-3)])  # 2e-5: fp32 tensor-core accumulation up to K=3200
@pytest.mark.parametrize("case", [
    dict(name="h2", k=5, cin=64, cout=128, stride=2, dil=1, tr=False, h=40, w=72, act=1, post=0, f32=False),
    dict(name="to_bn", k=5, cin=128, cout=33, stride=2, dil=1, tr=False, h=20, w=36, act=0, post=0, f32=True),
    dict(name="from_bn", k=3, cin=32, cout=128, stride=2, dil=1, tr=True, h=10, w=18, act=1, post=0, f32=False),
    dict(name="h12", k=5, cin=128, cout=64, stride=2, dil=1, tr=True, h=20, w=36, act=1, post=0, f32=False),
    dict(name="h13", k=5, cin=64, cout=3, stride=2, dil=1, tr=True, h=40, w=72, act=0, post=1, f32=True),
    dict(name="sinet_d1", k=3, cin=32, cout=32, stride=1, dil=1, tr=False, h=40, w=48, act=2, post=0, f32=False),
    dict(name="sinet_d16", k=3, cin=32, cout=32, stride=1, dil=16, tr=False, h=40, w=48, act=2, post=0, f32=False),
    dict(name="sinet_d128", k=3, cin=32, cout=32, stride=1, dil=128, tr=False, h=40, w=48, act=2, post=0, f32=False),
    dict(name="sinet_last", k=1, cin=32, cout=3, stride=1, dil=1, tr=False, h=24, w=40, act=0, post=2, f32=True),
], ids=lambda c: c["name"])
def test_generic_conv_tc_matches_float64(case, terms, tol):
    """Every layer shape that runs on the generic tcgen05 conv vs a float64 reference computed from the
    split-rounded values the kernel actually consumed."""
    from dsin_b200 import ops
    rng = np.random.default_rng(11)
    k, cin, cout = case["k"], case["cin"], case["cout"]
    x = rng.standard_normal((2, cin, case["h"], case["w"])).astype(np.float32)
    if case["tr"]:
        w_ref = (rng.standard_normal((k, k, cout, cin)) / np.sqrt(k * k * cin / 4)).astype(np.float32)
        w_pack = np.transpose(w_ref, (0, 1, 3, 2))
    else:
        w_ref = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
        w_pack = w_ref
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (0.3 * rng.standard_normal(cout)).astype(np.float32)
    layer = ops.ConvLayer(w_pack, scale, shift, stride=case["stride"], dilation=case["dil"], transposed=case["tr"],
                          act=case["act"], post=case["post"])
    tcl = ops.ConvTC(layer)
    xs = ops.f32_to_split(_nhwc(_dev(x)))
    xq = ops.split_to_f32(*xs).permute(0, 3, 1, 2).cpu().double()
    if case["tr"]:
        ref = O.conv2d_transpose_same_s2(xq, w_ref.astype(np.float64))
    else:
        ref = O.conv2d_same(xq, w_ref.astype(np.float64), stride=case["stride"], dilation=case["dil"])
    ref = ref * torch.tensor(scale).double().view(1, -1, 1, 1) + torch.tensor(shift).double().view(1, -1, 1, 1)
    if case["act"] == 1:
        ref = torch.relu(ref)
    elif case["act"] == 2:
        ref = torch.maximum(0.2 * ref, ref)
    res = None
    if not case["f32"]:
        r = rng.standard_normal(tuple(ref.shape)).astype(np.float32)
        res = ops.f32_to_split(_nhwc(_dev(r)))
        ref = ref + ops.split_to_f32(*res).permute(0, 3, 1, 2).cpu().double()
    if case["post"]:
        ref = O.denormalize(ref)
        if case["post"] == 1:
            ref = torch.clamp(ref, 0.0, 255.0)
    out = ops.conv_tc(xs, tcl, res1=res, terms=terms, out_f32=case["f32"])
    got = (out if case["f32"] else ops.split_to_f32(*out)).permute(0, 3, 1, 2).cpu().double()
    assert got.shape == ref.shape
    err = float((got - ref).abs().max())
    assert err < tol * max(1.0, float(ref.abs().max())), err


# ----------------------------------------------------------------------------- K9 MS-SSIM
def test_msssim_kernel_matches_reference_golden(golden_dir):
    """Device MS-SSIM (fp64) vs the values the reference's own numpy/scipy module produced
    (tests/golden/msssim_golden.npz), both call forms."""
    import os
    from dsin_b200 import utils
    g = np.load(os.path.join(golden_dir, "msssim_golden.npz"))
    for k in (0, 1):
        img, rec = g["img_%d" % k], g["rec_%d" % k].astype(np.float32)
        assert float(utils.msssim_standard(img, rec)) == pytest.approx(float(g["std_%d" % k]), abs=2e-7)
        assert float(utils.msssim_x_vs_rec(img, rec)) == pytest.approx(float(g["utils_%d" % k]), abs=2e-7)


def test_msssim_kernel_batch_matches_oracle_full_size():
    from dsin_b200 import ops
    x, y = synth.make_batch(2, 320, 1224, seed=77)
    rec = np.clip(x + np.random.default_rng(0).normal(0, 6, x.shape), 0, 255).astype(np.float32)
    got = ops.msssim(_nhwc(_dev(x)), _nhwc(_dev(rec)), form="standard")
    got2 = ops.msssim(_nhwc(_dev(x)), _nhwc(_dev(rec)), form="reference_call")
    for n in range(2):
        xi, ri = np.transpose(x[n], (1, 2, 0)), np.transpose(rec[n], (1, 2, 0))
        assert got[n] == pytest.approx(float(M.multi_scale_ssim(xi[None], ri[None])), abs=1e-9)
        assert got2[n] == pytest.approx(float(M.multi_scale_ssim(xi[..., None], ri[..., None])), abs=1e-9)


@pytest.mark.parametrize("mode", ["simt", "tc3"])
def test_probclass_modes_match_oracle(mode):
    """The tcgen05 probability model (the product path) and the all-CUDA-core kernel (cross-check, reached through
    the C ABI only from here) against the oracle."""
    from dsin_b200 import ops
    W = calibrated_weights(0)
    ae = make_ae(80, 144, W)
    rng = np.random.default_rng(6)
    c = W[O.ENC + "centers"]
    sym = torch.tensor(rng.integers(0, 6, (2, 32, 40, 153)))
    q = torch.tensor(c)[sym]
    ref = O.probclass_bitcost(q, sym, W)
    if mode == "tc3":
        bits = ae.pc_imgcomp.bitcost(q.cuda(), sym.cuda(), is_training=False, pad_value=float(c[0]))
        sums = bits._dsin_sum.cpu()
    else:
        bits, sums = ops.probclass_bits(q.cuda().contiguous(), sym.cuda().contiguous(), ae.pc_imgcomp.weights, float(c[0]))
        sums = sums.cpu()
    assert float((bits.cpu() - ref).abs().max()) < 1e-4
    assert torch.allclose(sums, ref.double().reshape(2, -1).sum(1), rtol=2e-7)


# ----------------------------------------------------------------------------- other BASELINE configs
def test_config4_geometry_320x960_encoder_quantizer_probclass():
    """BASELINE configs[3] geometry (320x960 crops): encoder + quantiser + probclass vs the oracle."""
    Wt = calibrated_weights(0)
    ae = make_ae(320, 960, Wt)
    x, _ = synth.make_batch(2, 320, 960, seed=4000)
    enc = ae.ae_imgcomp.encode(_dev(x))
    n_mism, bad, total = symbol_report(enc.symbols.cpu(), x, Wt)
    assert bad == 0 and n_mism <= max(2, total // 100000), (n_mism, bad, total)
    enc_ref = O.encode(torch.tensor(x), Wt)
    ref_bits = O.probclass_bitcost(enc_ref.qbar, enc_ref.symbols, Wt)
    bc = ae.pc_imgcomp.bitcost(enc_ref.qbar.cuda().contiguous(), enc_ref.symbols.cuda(), False,
                               pad_value=ae.pc_imgcomp.auto_pad_value(ae.ae_imgcomp))
    bpp_gpu = float(bc._dsin_sum.sum().item()) / (2 * 320 * 960)
    bpp_ref = float(O.bitcost_to_bpp(ref_bits, 2 * 320 * 960))
    assert abs(bpp_gpu - bpp_ref) <= 1e-5, (bpp_gpu, bpp_ref)


def test_config3_sifinder_isolated_batch():
    """BASELINE configs[2] (SI-Finder in isolation) at a reduced batch: every pair of the batch gets the
    same (row, col) as when it is processed alone (pairs are independent)."""
    from dsin_b200.siFinder import match_images
    x, y = _sif_case(320, 1224, 60, n=3)
    args = lambda sl: (_nhwc(_dev(x[sl])), _nhwc(_dev(y[sl])), _nhwc(_dev(y[sl])), 20, 24, True)  # noqa: E731
    _, _, _, row, col, best = match_images(*args(slice(0, 3)))
    for n in range(3):
        _, _, _, r1, c1, b1 = match_images(*args(slice(n, n + 1)))
        assert torch.equal(row[n], r1[0]) and torch.equal(col[n], c1[0]) and torch.equal(best[n], b1[0])


def test_ae_only_mode_and_error_paths():
    Wt = calibrated_weights(0)
    from parity_utils import configs
    from dsin_b200.AE import AE
    from dsin_b200.decoder_imgcomp import decoder
    from dsin_b200.encoder_imgcomp import encoder
    from dsin_b200.siFinder import siFinder
    from dsin_b200.siFull_img import SI_full_img
    from dsin_b200.siNet import siNet
    ae_config, pc_config = configs(80, 144)
    ae_config.AE_only = True
    ae = AE(ae_config, pc_config, encoder, decoder, siFinder, SI_full_img, siNet, "", weights=Wt)
    x, y = synth.make_batch(1, 80, 144, seed=9)
    y_dec, y_syn, x_dec, x_with_si, bpp = ae.siNet_get_reconstructed(x, y)
    assert float(np.abs(x_with_si).max()) == 0.0 and np.isfinite(x_dec).all() and bpp > 0
    with pytest.raises(NotImplementedError):
        ae.siNet_update(x, y)
    ae_config.arch = "nope"
    with pytest.raises(KeyError):
        AE(ae_config, pc_config, encoder, decoder, siFinder, SI_full_img, siNet, "", weights=Wt)


def test_cabi_error_codes_and_messages():
    """Bad arguments come back as negative codes with a message; nothing is launched."""
    import ctypes as C
    from dsin_b200 import _lib, ops
    h = ops.handle()
    lib = h.lib
    before = h.launch_count()
    d = _lib.ConvDesc(1, 8, 16, 3, 3, 3, 3, 3, 1, 0, 0, 0)  # stride 3 is not supported
    x = torch.zeros(1, 8, 16, 3, device="cuda")
    rc = lib.dsin_conv2d(h.ptr, C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), None, None, None,
                         None, C.c_void_p(x.data_ptr()), None)
    assert rc == -1 and b"stride" in lib.dsin_last_error(h.ptr)
    rc = lib.dsin_heatmap_quantize(h.ptr, None, None, 6, 1, 1, 1, 1, None, None, None, None, None, None, None)
    assert rc == -1
    rc = lib.dsin_sif_prepare(h.ptr, C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), 1, 30, 50, 20, 24,
                              C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()),
                              C.c_void_p(x.data_ptr()), None)
    assert rc == -1 and b"tile" in lib.dsin_last_error(h.ptr)
    assert h.launch_count() == before
    with pytest.raises(RuntimeError):
        h.check(rc)


def test_sinet_layer_forms_agree():
    """The three forms of the SI-Net's large-dilation layers -- row-band kernel (the product), tap streaming in the
    pixel-pair view (W/2 x 64 'channels', block-diagonal weights) and plain tap streaming -- against each other and
    against the oracle."""
    from dsin_b200 import siNet as sn
    Wt = calibrated_weights(0)
    ae = make_ae(120, 288, Wt)
    rng = np.random.default_rng(3)
    xd = np.clip(rng.normal(120, 50, (2, 3, 120, 288)), 0, 255).astype(np.float32)
    ys = np.clip(rng.normal(110, 60, (2, 3, 120, 288)), 0, 255).astype(np.float32)
    ref = O.denormalize(O.si_net(torch.cat([O.normalize(torch.tensor(xd)), O.normalize(torch.tensor(ys))], 1), Wt))
    outs = {}
    old = sn.BAND, sn.PAIR
    try:
        for form, (band, pair) in (("band", (True, True)), ("pair", (False, True)), ("plain", (False, False))):
            sn.BAND, sn.PAIR = band, pair
            outs[form] = ae._siNet.fused(_nhwc(_dev(xd)), _nhwc(_dev(ys))).cpu()
    finally:
        sn.BAND, sn.PAIR = old
    assert float((outs["band"] - outs["pair"]).abs().max()) < 2e-3
    assert float((outs["band"] - outs["plain"]).abs().max()) < 2e-3
    assert float((outs["band"] - ref).abs().max()) < 2e-2


def test_decode_side_region_equals_full_path():
    """The receiver-only region (qbar of x + y -> AE(y), decoder(x), SI-Finder, SI-Net) must give
    exactly what the full call gives: pairs and images are independent, so splitting the encoder pass
    changes nothing bit for bit."""
    ae = make_ae(80, 144, calibrated_weights(0))
    x, y = synth.make_batch(2, 80, 144, seed=31)
    xd, yd = _dev(x), _dev(y)
    full = ae.reconstruct_device(xd, yd)
    qb = full["qbar"].clone()
    keep = {k: full[k].clone() for k in ("y_dec", "x_dec", "y_syn", "x_with_si", "row", "col")}
    part = ae.decode_side_device(qb, yd)
    for k, v in keep.items():
        assert torch.equal(v, part[k]), k


def test_tf_checkpoint_save_restore_and_main_drop_in(tmp_path, monkeypatch):
    """SURVEY 8f N1+N2: `save_model` writes a TF-V2 checkpoint, `load_model` restores the scope-filtered
    variables from it, and the reference's README inference recipe (`python main.py` in a directory holding
    run_configs/, data_paths/<list>, weights/<name>/model.*) runs end to end: pair lists -> PNG decode ->
    centre crop -> siNet_get_reconstructed -> PNG named <i>_<bpp>bpp.png (uint8 truncation)."""
    from PIL import Image
    from dsin_b200 import main as dmain
    Wt = calibrated_weights(0)
    ae = make_ae(80, 144, Wt)
    prefix = str(tmp_path / "weights" / "tiny_model" / "model")
    ae.save_model(prefix)
    assert os.path.isfile(prefix + ".index") and os.path.isfile(prefix + ".data-00000-of-00001")
    # a fresh AE with different weights, then restore
    ae2 = make_ae(80, 144, synth.make_weights(3))
    ae2.load_model(prefix)
    x, y = synth.make_batch(1, 80, 144, seed=77)
    ref = ae.siNet_get_reconstructed(x, y)
    ref = [np.array(a) for a in ref]
    got = ae2.siNet_get_reconstructed(x, y)
    for a, b in zip(ref, got):
        assert np.array_equal(a, np.array(b))
    with pytest.raises(FileNotFoundError):
        ae2.load_model(str(tmp_path / "weights" / "missing" / "model"))

    # --- the drop-in directory layout of the reference's README recipe
    root = str(tmp_path / "kitti") + os.sep
    os.makedirs(root + "image_2"), os.makedirs(root + "image_3"), os.makedirs(tmp_path / "data_paths")
    xs, ys = synth.make_batch(2, 96, 160, seed=5)   # larger than the crop: exercises the centre crop
    lines = []
    for i in range(2):
        for cam, arr in (("image_2", xs), ("image_3", ys)):
            rel = "%s/%06d_10.png" % (cam, i)
            Image.fromarray(arr[i].transpose(1, 2, 0).astype(np.uint8), "RGB").save(root + rel)
            lines.append(rel)
    (tmp_path / "data_paths" / "KITTI_stereo_test.txt").write_text("\n".join(lines) + "\n")
    cfg_src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dsin_b200", "run_configs")
    txt = open(os.path.join(cfg_src, "ae_run_configs")).read()
    txt = txt.replace("crop_size = (320,1224)", "crop_size = (80,144)")
    txt = txt.replace("root_data = '/root/sharedfolder2/'", "root_data = %r" % root)
    txt = txt.replace("load_model_name = 'KITTI_stereo_target_bpp0.02'", "load_model_name = 'tiny_model'")
    assert "tiny_model" in txt and "(80,144)" in txt and root in txt
    os.makedirs(tmp_path / "run_configs")
    (tmp_path / "run_configs" / "ae_run_configs").write_text(txt)
    monkeypatch.chdir(tmp_path)
    args = dmain.build_parser().parse_args(
        ["-ae_config", str(tmp_path / "run_configs" / "ae_run_configs"), "--create_loss_list"])
    bpps = dmain.main(dmain.get_run_params(args), args)
    assert len(bpps) == 2 and all(b > 0 for b in bpps)
    # same numbers as the direct call on the centre crops
    for i in range(2):
        xc = xs[i:i + 1, :, 8:88, 8:152].astype(np.uint8)
        yc = ys[i:i + 1, :, 8:88, 8:152].astype(np.uint8)
        _yd, _ys, _xd, xsi, bpp = ae.siNet_get_reconstructed(xc, yc)
        assert float(bpp) == bpps[i]
        png = str(tmp_path / "images" / "tiny_model" / ("%d_%.5fbpp.png" % (i, bpp)))
        assert os.path.isfile(png), os.listdir(tmp_path / "images" / "tiny_model")
        assert np.array_equal(np.asarray(Image.open(png)),
                              np.clip(np.array(xsi[0]), 0, 255).transpose(1, 2, 0).astype("uint8"))
    for name in ("bpp_list_", "l1_list_", "psnr_list_", "msssim_list_", "mse_list_x_y_syn_", "avg_Pearson_list_x_y_syn_"):
        assert len(open(str(tmp_path / "images" / (name + "tiny_model.txt"))).read().split()) == 2


def test_cuda_graph_replay_equals_eager_and_follows_weight_reload():
    """The numpy entry point replays two captured CUDA graphs; results must be bit-identical to the eager
    launch sequence, for changing inputs, and a weight reload must invalidate the captured graphs."""
    ae = make_ae(80, 144, calibrated_weights(0))
    assert ae.use_cuda_graph
    pairs = [synth.make_batch(2, 80, 144, seed=s) for s in (3, 4)]
    got = []
    for x, y in pairs:                      # first call captures, second replays with new inputs
        got.append([np.array(a) for a in ae.siNet_get_reconstructed(x.astype(np.uint8), y.astype(np.uint8))])
    assert len(ae._graphs) == 1
    ae.use_cuda_graph = False
    for (x, y), g in zip(pairs, got):
        ref = ae.siNet_get_reconstructed(x.astype(np.uint8), y.astype(np.uint8))
        for a, b in zip(ref, g):
            assert np.array_equal(np.array(a), b)
    ae.use_cuda_graph = True
    ae.set_weights(calibrated_weights(1))
    assert not ae._graphs
    x, y = pairs[0]
    new = [np.array(a) for a in ae.siNet_get_reconstructed(x, y)]
    assert not np.array_equal(new[2], got[0][2])          # different weights -> different x_dec
    ae.use_cuda_graph = False
    ref = ae.siNet_get_reconstructed(x, y)
    for a, b in zip(ref, new):
        assert np.array_equal(np.array(a), b)


# ----------------------------------------------------------------------------- kernels vs the reference's own functions
def test_kernels_match_reference_elementwise_functions(golden_dir):
    """The CUDA kernels against outputs of the reference's OWN Python (tests/golden/tf_pieces_golden.npz, produced by
    running the ast-extracted functions under a numpy stand-in for their elementwise tf ops): SI-Finder normalisation +
    colour transform (src/siFinder.py:56-73,138-154), heatmap and masked bottleneck
    (src/autoencoder_imgcomp.py:173-201), quantiser (src/quantizer_imgcomp.py:43-100)."""
    from dsin_b200 import ops
    g = np.load(os.path.join(golden_dir, "tf_pieces_golden.npz"))
    # K5: five 20x24 patches side by side form a 20x120 image; q = transformed patches, r = transformed image
    patches = g["sif_in"]                                              # (5,20,24,3)
    img = np.ascontiguousarray(np.concatenate(list(patches), axis=1))[None]  # (1,20,120,3)
    q, r, _ps, _ys = ops.sif_prepare(_dev(img), _dev(img), 20, 24)
    assert torch.equal(q.cpu().reshape(5, 20, 24, 3), torch.tensor(g["sif_rgb"]))
    assert torch.equal(r.cpu()[0], torch.tensor(np.concatenate(list(g["sif_rgb"]), axis=1)))
    # K3 heatmap + mask
    z33 = g["z33"]
    c = g["q_centers"]
    out = ops.heatmap_quantize(_nhwc(_dev(z33)), _dev(c), full=True)
    hm, z = out[5].cpu(), out[4].cpu()
    assert float((hm - torch.tensor(g["heatmap3d"])).abs().max()) <= 4e-6
    assert float((z - torch.tensor(g["z_masked"])).abs().max()) <= 2e-5
    # K3 quantiser on the golden input: heatmap channel +30 -> H3D == 1 everywhere, so z == q_in exactly
    z33q = np.concatenate([np.full((2, 1, 6, 7), 30.0, np.float32), g["q_in"]], axis=1)
    _qn, qbar, sym, qhard, zz, hm1 = ops.heatmap_quantize(_nhwc(_dev(z33q)), _dev(c), full=True)
    assert torch.equal(hm1.cpu(), torch.ones(2, 32, 6, 7)) and torch.equal(zz.cpu(), torch.tensor(g["q_in"]))
    assert torch.equal(sym.cpu(), torch.tensor(g["q_symbols"]))        # incl. on-centre and midway inputs
    assert torch.equal(qhard.cpu(), torch.tensor(g["q_hard"]))
    qsoft_ref, qhard_ref = torch.tensor(g["q_soft"]), torch.tensor(g["q_hard"])
    assert float((qbar.cpu() - (qsoft_ref + (qhard_ref - qsoft_ref))).abs().max()) <= 6e-7


def test_sender_receiver_round_trip_matches_the_one_call_path():
    """compress(x) -> bytes -> decompress(bytes, y): the receiver feeds the decoder qhard = centres[symbols] while the
    one-call path (and the reference, src/AE.py:56) feeds qbar = qsoft + (qhard - qsoft), one fp32 rounding away.
    The gap must stay at rounding level: same symbols, same (row, col), images within 0.05 grey levels."""
    ae = make_ae(80, 144, calibrated_weights(0))
    x, y = synth.make_batch(2, 80, 144, seed=55)
    x8, y8 = x.astype(np.uint8), y.astype(np.uint8)
    one = [np.array(a) for a in ae.siNet_get_reconstructed(x8, y8)]
    sym_one = ae.last["symbols"].clone()
    row_one, col_one = ae.last["row"].clone(), ae.last["col"].clone()
    blobs = ae.compress(x8)
    y_dec, y_syn, x_dec, x_with_si = [np.array(a) for a in ae.decompress(blobs, y8)]
    assert torch.equal(ae.last["symbols"], sym_one)
    assert torch.equal(ae.last["row"], row_one) and torch.equal(ae.last["col"], col_one)
    assert np.array_equal(y_dec, one[0]) and np.array_equal(y_syn, one[1])
    assert float(np.abs(x_dec - one[2]).max()) < 0.05 and float(np.abs(x_with_si - one[3]).max()) < 0.05
    with pytest.raises(ValueError):  # a container of another geometry is refused from its header
        make_ae(160, 144, calibrated_weights(0)).decompress(blobs, np.zeros((2, 3, 160, 144), np.uint8))


def test_load_model_variable_selection_follows_the_reference(tmp_path):
    """src/AE.py:158-175: the SI-Net scope is restored when load_train_step, or when testing without training; a
    first SI training run from an AE-only checkpoint (train_model) restores the AE scopes and keeps the SI-Net's
    initialiser values."""
    from dsin_b200 import tf_checkpoint
    Wt = calibrated_weights(0)
    ae_only = {k: v for k, v in Wt.items() if not k.startswith("siNetwork/")}
    prefix = str(tmp_path / "ae_only" / "model")
    tf_checkpoint.write_checkpoint(prefix, ae_only)
    ae = make_ae(80, 144, synth.make_weights(5))
    with pytest.raises(KeyError):          # test_model=True, train_model=False: siNetwork/* is part of the restore list
        ae.load_model(prefix)
    ae.ae_config.train_model, ae.ae_config.test_model = True, False
    si_before = {k: v.copy() for k, v in ae.weights.items() if k.startswith("siNetwork/")}
    ae.load_model(prefix)
    for k, v in si_before.items():
        assert np.array_equal(ae.weights[k], v)                      # SI-Net kept
    assert np.array_equal(ae.weights[O.ENC + "h1/weights"], Wt[O.ENC + "h1/weights"])  # AE restored


def test_sharded_batch_equals_whole_batch_bit_for_bit():
    """BASELINE configs[4] / SURVEY 8e: the same global batch on 1 and on 2 GPUs gives bit-identical per-image outputs.
    Needs two GPUs (gpurun --gpus 2); with one GPU the single-process half of the script still checks that batch
    composition (micro-batch size 3 vs the whole shard) changes nothing."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "multi_gpu_equivalence.py")
    n = min(2, torch.cuda.device_count())
    if n >= 2:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29631", script, "--global-batch", "6", "--hw", "160x288"]
    else:
        cmd = [sys.executable, script, "--global-batch", "5", "--hw", "160x288"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    assert '"identical_per_image": true' in line, line


@pytest.mark.parametrize("geom", [(80, 144, 4, 12), (320, 1224, 8, 16)])
def test_pipelined_numpy_call_equals_single_shot_and_per_chunk_calls(geom):
    """A numpy call with more pairs than AE.e2e_chunk runs as a pipeline of chunks over three streams (H2D, kernels,
    D2H).  Pairs are independent, so it must return exactly what the same pairs give chunk by chunk, and `last` must
    carry the whole batch; repeated calls (recycled pinned buffers, static graph buffers, events) must stay identical.
    Against the one-graph run of the whole batch only the integers are compared bit for bit (a 12-pair launch tiles
    the same arithmetic, but that is the property test_gpu_freerun / multi_gpu_equivalence hold, not this test)."""
    H, W, c, B = geom  # the full-size case has D2H copies long enough to overlap the next chunk's first kernels
    ae = make_ae(H, W, calibrated_weights(0))
    ae.e2e_chunk = c
    x, y = synth.make_batch(B, H, W, seed=77)
    x8, y8 = x.astype(np.uint8), y.astype(np.uint8)
    runs = []
    for _ in range(3):
        res = [np.array(a) for a in ae.siNet_get_reconstructed(x8, y8)]
        runs.append((res, {k: v.clone() for k, v in ae.last.items()}))
    for res, last in runs[1:]:
        for a, b in zip(res, runs[0][0]):
            assert np.array_equal(a, b)
        for k in last:
            assert torch.equal(last[k], runs[0][1][k]), k
    res, last = runs[0]
    assert last["symbols"].shape[0] == B and last["row"].shape[0] == B and last["bits_sum"].shape[0] == B
    bits_total = 0.0
    for k in range(B // c):  # chunk by chunk through the single-shot path (c pairs <= e2e_chunk)
        sl = slice(c * k, c * k + c)
        part = [np.array(a) for a in ae.siNet_get_reconstructed(x8[sl], y8[sl])]
        for i in range(4):
            assert np.array_equal(part[i], res[i][sl]), (k, i)
        assert torch.equal(ae.last["symbols"], last["symbols"][sl]) and torch.equal(ae.last["row"], last["row"][sl])
        bits_total += float(ae.last["bits_sum"].sum().item())
    assert abs(float(res[4]) - bits_total / (B * H * W)) <= 1e-6
    ae.e2e_chunk = None  # the whole batch as one launch sequence
    whole = [np.array(a) for a in ae.siNet_get_reconstructed(x8, y8)]
    assert torch.equal(ae.last["symbols"], last["symbols"]) and torch.equal(ae.last["row"], last["row"])
    assert torch.equal(ae.last["col"], last["col"])
    for i in range(4):
        assert float(np.abs(whole[i] - res[i]).max()) < 1e-2
    assert abs(float(whole[4]) - float(res[4])) <= 1e-6
