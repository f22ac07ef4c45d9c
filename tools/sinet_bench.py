"""A/B timing of the SI-Net (src/siNet.py:29-41) at batch B on one box: row-band kernel for the large dilations
(siNet.BAND, default) vs the tap-streaming pixel-pair form, both orders; prints ms per SI-Net pass and per layer."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_utils import make_ae  # noqa: E402
from dsin_b200 import ops, siNet as sn, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    H, W = 320, 1224
    ae = make_ae(H, W, synth.make_weights(0, residual_gamma=0.25))
    a = torch.rand(B, H, W, 3, device="cuda") * 255
    b = torch.rand(B, H, W, 3, device="cuda") * 255
    outs = {}
    for band in (True, False, False, True):
        sn.BAND = band
        for _ in range(3):
            y = ae._siNet.fused(a, b, terms=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y = ae._siNet.fused(a, b, terms=3)
        e1.record()
        torch.cuda.synchronize()
        outs[band] = y._dsin_nhwc.clone()
        print("BAND=%s: %.3f ms per SI-Net pass at batch %d" % (band, e0.elapsed_time(e1) / reps, B), flush=True)
    print("max |band - pair| = %.3e grey levels" % float((outs[True] - outs[False]).abs().max()))
    # per layer
    cur = ops.f32_to_split(torch.randn(B, H, W, 32, device="cuda"))
    for li in (3, 4, 5, 6, 7):
        rate = sn.SiNet.RATES[li]
        tcl = ae._siNet._tc[li - 1]
        for name, flags in (("band", 0), ("stream", ops.CONV_NO_HALO)):
            for _ in range(2):
                ops.conv_tc(cur, tcl, terms=3, flags=flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.conv_tc(cur, tcl, terms=3, flags=flags)
            e1.record()
            torch.cuda.synchronize()
            print("  rate %3d %-6s %.3f ms" % (rate, name, e0.elapsed_time(e1) / reps), flush=True)


if __name__ == "__main__":
    main()
