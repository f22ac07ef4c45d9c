"""Runs the non-trunk kernels of one batch-8 step once inside a cudaProfilerStart/Stop range, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/<name> \
        python tools/ncu_targets.py [--what sinet,probclass,sif,quant,trunk] [--policy mixed]
Every kernel is warmed up once before the profiled range."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_utils import make_ae  # noqa: E402
from dsin_b200 import ops, precision, synth  # noqa: E402
from dsin_b200.siFinder import match_images  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="sinet,probclass,quant,sif")
    ap.add_argument("--policy", default=None)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    what = args.what.split(",")
    torch.cuda.set_device(0)
    B, H, W = args.batch, 320, 1224
    pol = precision.get(args.policy)
    ae = make_ae(H, W, synth.make_weights(0, residual_gamma=0.25), precision=pol)
    x, y = synth.make_batch(B, H, W, seed=1000)
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    out = ae.reconstruct_device(xd, yd)  # warm-up of everything + realistic inputs for the pieces below
    torch.cuda.synchronize()
    x_dec, y_syn = out["x_dec"]._dsin_nhwc, out["y_syn"]._dsin_nhwc
    y_dec = out["y_dec"]._dsin_nhwc
    y_nhwc = ops.nchw_to_nhwc(yd)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if "step" in what:    # one whole eager step (every kernel of the timed path, in order)
        ae.reconstruct_device(xd, yd)
    if "sinet" in what:
        ae._siNet.fused(x_dec, y_syn, terms=pol.sinet)
    if "probclass" in what:
        ae.pc_imgcomp.bitcost(out["qbar"], out["symbols"], False, pad_value=ae.pc_imgcomp.auto_pad_value(ae.ae_imgcomp),
                              terms=pol.probclass)
    if "quant" in what:
        z33 = torch.randn(2 * B, 40, 153, 33, device="cuda")
        ops.heatmap_quantize(z33, ae.ae_imgcomp._centers, full=True)
    if "sif" in what:
        match_images(x_dec, y_nhwc, y_dec, 20, 24, True)
    if "trunk" in what:   # decoder of 2B images: the fp16-operand trunk (conv_ws) + from_bn / h12 / h13
        ae.ae_imgcomp.decode(torch.cat([out["qbar"], out["qbar"]]), terms=pol.dec)
    if "enc" in what:     # encoder of 2B images: the fp32-class trunk (conv_h3) + h1 / h2 / to_bn + quantiser
        ae.ae_imgcomp.encode(torch.cat([yd, xd]), terms=pol.enc_x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
