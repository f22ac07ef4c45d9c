"""Times one trunk layer (3x3 128->128, 16 images of 80x306) of the tcgen05 conv with 0 / 2 residual inputs and
3 / 1 MMA terms; DSIN_NO_CTA2=1 selects the one-CTA kernel instead of the CTA-pair kernel."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from dsin_b200 import ops
n, hh, ww = 16, 80, 306
rng = np.random.default_rng(0)
w = (rng.standard_normal((3, 3, 128, 128)) / 34).astype(np.float32)
layer = ops.ConvLayer(w, np.ones(128, np.float32), np.zeros(128, np.float32), act=ops.ACT_RELU)
tcl = ops.ConvTC(layer)
x = ops.f32_to_split(torch.randn(n, hh, ww, 128, device="cuda"))
r = ops.f32_to_split(torch.randn(n, hh, ww, 128, device="cuda"))
for terms in (3, 1):
    for res in (False, True):
        for _ in range(3):
            ops.conv_tc(x, tcl, res1=r if res else None, res2=r if res else None, terms=terms)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_tc(x, tcl, res1=r if res else None, res2=r if res else None, terms=terms)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print("cta_pairs=%s terms=%d residuals=%d: %.1f us  (%.0f TFLOP/s algorithmic)" % (
            "DSIN_NO_CTA2" not in os.environ, terms, 2 * res, us, 2.0 * n * hh * ww * 9 * 128 * 128 / us / 1e6))
