"""The loss arithmetic of the validation pass (/root/reference/src/Distortions_imgcomp.py, src/AE.py:76-99).

Only what `AE.siNet_validate` needs (forward, inference mode): the distortion that is minimised, `get_loss` and the
regularisation terms.  The reductions over images and over the bottleneck run on the GPU (csrc/loss.cu,
`ops.validation_terms`, fp64 per-image sums); what is left here is scalar float32 arithmetic in the reference's order.
Gradients, optimisers and the learning-rate schedule (src/training_helpers_imgcomp.py) are not built."""
from __future__ import annotations

import re

import numpy as np

f32 = np.float32


def regularization_loss(variables, scope, factor, factor_centers=None):
    """`tf.losses.get_regularization_loss(scope)` over a dict of TF variables (name -> array).

    slim registers `factor * l2_loss(w)` = factor * sum(w^2) / 2 for every conv / deconv `weights` variable created
    inside `_building_ctx` (src/autoencoder_imgcomp.py:98-104) and `factor_centers * l2_loss(centers)` for the centres
    (src/quantizer_imgcomp.py:18-24); the op names start with the variable's full name.  The collection is then
    filtered with `re.match(scope, op.name)` -- a PREFIX match.  In the graph src/AE.py:51-56 builds, the variables live
    under 'encoder/encoder_body/...' and 'decoder/...' while the scopes asked for are 'autoencoder/encoder' and
    'autoencoder/decoder' (src/autoencoder_imgcomp.py:21-23,80-86), so nothing matches and the reference's
    regularisation term is 0.0.  The rule is implemented in full so that differently scoped weights behave as TF would."""
    rx = re.compile(scope)
    total = 0.0
    for name in sorted(variables):
        if not rx.match(name):
            continue
        if name.endswith("/weights") and factor:
            w = np.asarray(variables[name], dtype=np.float64)
            total += float(factor) * 0.5 * float(np.sum(w * w))
        elif name.endswith("/centers") and factor_centers:
            w = np.asarray(variables[name], dtype=np.float64)
            total += float(factor_centers) * 0.5 * float(np.sum(w * w))
    return f32(total)


def distortion_to_minimize(config, dist_sums, img_elems):
    """`Distortions(config, x, x_out, is_training=True).d_loss_scaled` (src/Distortions_imgcomp.py:7-56) from the
    per-image sums of |x_out - x| (mae) or (x_out - x)^2 (mse, psnr): per-image mean, then the mean over the batch.
    With is_training=True -- how src/AE.py:79 builds it, also for validation -- the minimised distortion is computed on
    the float images, not on integer casts (src/Distortions_imgcomp.py:19-21)."""
    kind = config.distortion_to_minimize
    per_img = (np.asarray(dist_sums, dtype=np.float64) / float(img_elems)).astype(f32)
    if kind in ("mae", "mse"):
        return f32(np.mean(per_img, dtype=f32))
    if kind == "psnr":
        psnr = (f32(10.0) * (np.log(f32(255.0 * 255.0) / per_img) / f32(np.log(10.0)))).astype(f32)
        return f32(f32(config.K_psnr) - np.mean(psnr, dtype=f32))
    raise NotImplementedError("distortion_to_minimize = %r: the MS-SSIM training loss is not built "
                              "(the shipped config minimises mae, src/run_configs/ae_run_configs:23)" % (kind,))


def squared_distortion(config):
    """Whether the per-image reduction needed by `distortion_to_minimize` is a sum of squares."""
    return config.distortion_to_minimize in ("mse", "psnr")


def get_loss(config, ae, pc, d_loss_scaled, H_real, H_mask):
    """src/Distortions_imgcomp.py:113-146 with the two reductions (`H_real` = mean(bc), `H_mask` = mean(bc * heatmap))
    already done -> (total_loss, H_real, pc_comps, ae_comps), float32 scalars."""
    assert config.H_target
    H_real, H_mask = f32(H_real), f32(H_mask)
    H_soft = f32(f32(0.5) * f32(H_mask + H_real))
    pc_loss = f32(f32(config.beta) * np.maximum(f32(H_soft - f32(config.H_target)), f32(0)))
    reg_probclass = pc.regularization_loss()
    if reg_probclass is None:
        reg_probclass = f32(0)
    reg_enc = ae.encoder_regularization_loss()
    reg_dec = ae.decoder_regularization_loss()
    reg_loss = f32(f32(reg_probclass + reg_enc) + reg_dec)
    pc_comps = [("H_mask", H_mask), ("H_real", H_real), ("pc_loss", pc_loss), ("reg", reg_probclass)]
    ae_comps = [("d_loss_scaled", f32(d_loss_scaled)), ("reg_enc_dec", f32(reg_enc + reg_dec))]
    total_loss = f32(f32(f32(d_loss_scaled) + pc_loss) + reg_loss)
    return total_loss, H_real, pc_comps, ae_comps
