"""Diagnostic for csrc/conv_ws.cu: one-tap identity filters show which input pixel/channel each output reads."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsin_b200 import ops  # noqa: E402

torch.cuda.set_device(0)
n, hh, ww = 1, 32, 24
# every (pixel, channel) gets a unique fp16-exact code: pixel index in the integer part (< 2048), channel / 128 fraction
pix = torch.arange(hh * ww, dtype=torch.float32).view(1, hh, ww, 1)
ch = torch.arange(128, dtype=torch.float32).view(1, 1, 1, 128)
x = (pix + 0 * ch).cuda()
xh = x.half().contiguous()           # value = pixel index (exact up to 2048)
xc = (ch + 0 * pix).cuda().half().contiguous()   # value = channel index
for (ky, kx) in [(1, 1), (0, 0), (1, 2), (2, 1), (1, 0), (2, 2)]:
    w = np.zeros((3, 3, 128, 128), np.float32)
    for c in range(128):
        w[ky, kx, c, c] = 1.0
    layer = ops.ConvLayer(w, np.ones(128, np.float32), np.zeros(128, np.float32), act=ops.ACT_NONE)
    tcl = ops.ConvTC(layer)
    for flags, name in ((0, "weight_stationary"), (ops.CONV_NO_WEIGHT_STATIONARY, "streaming")):
        y, _ = ops.conv_tc((xh, None), tcl, terms=1, flags=flags)
        yc, _ = ops.conv_tc((xc, None), tcl, terms=1, flags=flags)
        torch.cuda.synchronize()
        y = y.float().cpu()[0]
        yc = yc.float().cpu()[0]
        # expected: y[oy, ox, c] = pixel index of (oy + ky - 1, ox + kx - 1) or 0 outside
        exp = torch.zeros(hh, ww)
        for oy in range(hh):
            for ox in range(ww):
                iy, ix = oy + ky - 1, ox + kx - 1
                if 0 <= iy < hh and 0 <= ix < ww:
                    exp[oy, ox] = iy * ww + ix
        ok_pix = (y[:, :, 0] == exp)
        ok_allc = (y == exp.unsqueeze(-1)).all(-1)
        okc = (yc == ch[0, 0].expand(hh, ww, 128))
        inside = exp > 0
        print("tap (%d,%d) %-15s pixel-source correct (ch0) %.3f  all channels %.3f | channel identity correct %.3f"
              % (ky, kx, name, float(ok_pix.float().mean()), float(ok_allc.float().mean()),
                 float(okc[inside].float().mean())))
        if name != "streaming" and float(ok_allc.float().mean()) < 0.99:
            for oy, ox in [(0, 0), (0, 1), (0, 7), (1, 0), (5, 3), (17, 9), (16, 8)]:
                got = y[oy, ox]
                print("    out(%2d,%2d) expects %5.0f; got ch0..3 %s ch8..11 %s ch64..67 %s | chan-id ch0..3 %s ch64..66 %s"
                      % (oy, ox, exp[oy, ox], got[0:4].tolist(), got[8:12].tolist(), got[64:68].tolist(),
                         yc[oy, ox, 0:4].tolist(), yc[oy, ox, 64:67].tolist()))
