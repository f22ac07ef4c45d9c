"""AE.siNet_validate (the forward half of SURVEY 8f N4) against the oracle's restatement of loss_test
(src/AE.py:76-99,120-131).  The oracle's loss arithmetic itself is pinned to the reference's own Distortions class and
get_loss in tests/test_oracle_golden.py (tests/golden/loss_pieces_golden.npz).

Tolerances (floating point, so stated here): the oracle is replayed with the GPU's float64-adjudicated symbol
tie-breaks forced (as in test_gpu_freerun.py); against that replay
  * H_real, H_mask (bits per symbol):  |d| <= 1e-5           -- same bar as bpp
  * loss_test:                         |d| <= 1e-5 * |loss|  -- the rate term is beta = 500 times H_soft
  * the two image terms (grey levels): |d| <= 2e-3           -- the shipped policy runs the decoders on fp16 operands
                                                                (0.07 grey levels rms on x_dec, averaging out in a mean)
"""
import numpy as np
import pytest
import torch

from oracle import dsin_oracle as O

import oracle_cache
from parity_utils import make_ae
from test_gpu_freerun import _forced

pytestmark = pytest.mark.gpu


def _oracle_loss(ae, x, y, Wt, fx=None, fy=None):
    c = ae.ae_config
    return O.validation_loss(x.astype(np.float32), y.astype(np.float32), Wt, si_weight=c.si_weight, beta=c.beta,
                             H_target=c.H_target, distortion=c.distortion_to_minimize, K_psnr=c.K_psnr,
                             regularization_factor=c.regularization_factor,
                             regularization_factor_centers=c.regularization_factor_centers,
                             force_symbols_x=None if fx is None else torch.as_tensor(fx),
                             force_symbols_y=None if fy is None else torch.as_tensor(fy))


def _check(case_name):
    Wt, d = oracle_cache.get(case_name)
    c = oracle_cache.CASES[case_name]
    ae = make_ae(c["H"], c["W"], Wt)
    x, y = d["x"], d["y"]
    loss = ae.siNet_validate(x, y)
    comps = ae.last_loss
    fx, nx = _forced(ae.last["symbols"].cpu().numpy(), d["sym32_x"], d["margin64_x"], "x")
    fy, ny = _forced(ae.last["symbols_y"].cpu().numpy(), d["sym32_y"], d["margin64_y"], "y")
    ref = _oracle_loss(ae, x, y, Wt, fx, fy)
    print("%s: loss_test gpu %.6f oracle %.6f (tie-breaks x %d y %d); H_real %.7f / %.7f, H_mask %.7f / %.7f, "
          "d_loss %.5f / %.5f, loss_siNet %.5f / %.5f" % (case_name, loss, ref.loss, nx, ny, comps["H_real"], ref.H_real,
                                                         comps["H_mask"], ref.H_mask, comps["d_loss"], ref.d_loss_scaled,
                                                         comps["loss_siNet"], ref.loss_siNet))
    assert abs(float(comps["H_real"]) - ref.H_real) <= 1e-5
    assert abs(float(comps["H_mask"]) - ref.H_mask) <= 1e-5
    assert abs(float(comps["d_loss"]) - ref.d_loss_scaled) <= 2e-3
    assert abs(float(comps["loss_siNet"]) - ref.loss_siNet) <= 2e-3
    assert abs(float(comps["pc_loss"]) - ref.pc_loss) <= 1e-5 * abs(ref.pc_loss) + 1e-6
    assert float(comps["reg_enc_dec"]) == ref.reg_loss == 0.0  # the reference's scope filter matches nothing
    assert abs(loss - ref.loss) <= 1e-5 * abs(ref.loss)
    return ae, x, y, loss


def test_validation_loss_small_batch2():
    ae, x, y, loss = _check("small")
    # eager launches and CUDA-graph replays are the same computation
    ae.use_cuda_graph = False
    assert ae.siNet_validate(x, y) == loss
    # the loss of a batch: bit sums and image sums add up over the images
    la = ae.siNet_validate(x[:1], y[:1])
    ca = dict(ae.last_loss)
    lb = ae.siNet_validate(x[1:], y[1:])
    cb = dict(ae.last_loss)
    ae.siNet_validate(x, y)
    for k in ("H_real", "H_mask", "d_loss", "loss_siNet"):
        assert abs(0.5 * (float(ca[k]) + float(cb[k])) - float(ae.last_loss[k])) <= 2e-6 * max(1.0, abs(float(ca[k]))), k
    assert la != lb


def test_validation_loss_full_size():
    _check("full1")


def test_validation_loss_other_distortions_and_terms():
    """mse / psnr distortions, the clamped rate term (H_target above H_soft) and AE_only, against the same arithmetic
    evaluated in float64 from the tensors of the pass itself."""
    Wt, d = oracle_cache.get("small")
    c = oracle_cache.CASES["small"]
    ae = make_ae(c["H"], c["W"], Wt)
    x, y = d["x"], d["y"]
    xf = torch.as_tensor(x.astype(np.float64))
    for kind, h_target in (("mse", 0.04), ("psnr", 0.04), ("mae", 50.0)):
        ae.ae_config.distortion_to_minimize = kind
        ae.ae_config.H_target = h_target
        loss = ae.siNet_validate(x, y)
        out = ae.last
        x_dec, x_si = out["x_dec"].double().cpu(), out["x_with_si"].double().cpu()
        bc, hm = out["bits"].double().cpu(), out["heatmap"].double().cpu()
        dist = O.distortion_to_minimize(xf, x_dec, kind, ae.ae_config.K_psnr)
        total, _hr, _hm, pc_loss = O.get_loss((1.0 - ae.si_weight) * dist, bc, hm, ae.ae_config.beta, h_target)
        want = float(total + ae.si_weight * (xf - x_si).abs().mean())
        print(kind, h_target, loss, want)
        assert abs(loss - want) <= 2e-6 * abs(want)
        if h_target == 50.0:
            assert float(pc_loss) == 0.0 and float(ae.last_loss["pc_loss"]) == 0.0
    with pytest.raises(NotImplementedError):
        ae.ae_config.distortion_to_minimize = "ms_ssim"
        ae.siNet_validate(x, y)


def test_validation_loss_ae_only():
    from dsin_b200.AE import AE
    from dsin_b200.decoder_imgcomp import decoder
    from dsin_b200.encoder_imgcomp import encoder
    from dsin_b200.siFinder import siFinder
    from dsin_b200.siFull_img import SI_full_img
    from dsin_b200.siNet import siNet
    from parity_utils import CFG, configs
    Wt, d = oracle_cache.get("small")
    c = oracle_cache.CASES["small"]
    ae_config, pc_config = configs(c["H"], c["W"])
    ae_config.AE_only = True
    ae = AE(ae_config, pc_config, encoder, decoder, siFinder, SI_full_img, siNet, CFG, weights=Wt)
    x, y = d["x"], d["y"]
    loss = ae.siNet_validate(x, y)
    cm = ae.last_loss
    # si_weight = 0 (src/AE.py:19): loss = d_loss + pc_loss, no SI term
    assert float(cm["loss_siNet"]) == 0.0
    assert abs(loss - (float(cm["d_loss"]) + float(cm["pc_loss"]))) <= 1e-6 * abs(loss)
    full = make_ae(c["H"], c["W"], Wt)
    full.siNet_validate(x, y)
    assert float(full.last_loss["d_loss"]) == float(cm["d_loss"]) and float(full.last_loss["H_real"]) == float(cm["H_real"])
