"""A/B timing of the trunk-layer kernels (3x3 128->128) on one box: a whole 32-layer trunk pass with the real residual
pattern (conv1: none, conv2: one, every third conv2: two), repeated long enough to reach the power-capped steady state.
  terms 3: halo kernel conv_h3 (default) vs tap-streaming conv_tc2<3> (CONV_NO_HALO)
  terms 1: weight-stationary conv_ws (default) vs tap-streaming conv_tc2<1> (CONV_NO_WEIGHT_STATIONARY)
Prints ms per trunk pass and algorithmic TFLOP/s, plus the SM clock sampled during each run."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsin_b200 import ops  # noqa: E402


class Clock(object):
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.n, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.v, self.p, self.stop = [], [], False

    def run(self):
        while not self.stop:
            self.v.append(self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM))
            self.p.append(self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            time.sleep(0.01)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hh, ww = 80, 306
    rng = np.random.default_rng(0)
    layers = []
    for _ in range(4):
        w = (rng.standard_normal((3, 3, 128, 128)) / 34).astype(np.float32)
        layers.append(ops.ConvTC(ops.ConvLayer(w, np.ones(128, np.float32), np.zeros(128, np.float32), act=ops.ACT_RELU)))
    x3 = ops.f32_to_split(torch.randn(n, hh, ww, 128, device="cuda"))
    flop = 32 * 2.0 * n * hh * ww * 9 * 128 * 128

    def trunk(cur, terms, flags):
        r0 = cur
        for b in range(5):
            rb = cur
            for i in range(3):
                t = ops.conv_tc(cur, layers[0], terms=terms, flags=flags)
                cur = ops.conv_tc(t, layers[1], res1=cur, res2=rb if i == 2 else None, terms=terms, flags=flags)
        t = ops.conv_tc(cur, layers[2], terms=terms, flags=flags)
        return ops.conv_tc(t, layers[3], res1=cur, res2=r0, terms=terms, flags=flags)

    order = ((3, (("halo (conv_h3)", 0), ("streaming (conv_tc2<3>)", ops.CONV_NO_HALO))),
             (1, (("weight-stationary (conv_ws)", 0), ("streaming (conv_tc2<1>)", ops.CONV_NO_WEIGHT_STATIONARY))))
    order = order + tuple((t, v[::-1]) for t, v in order)  # A B, then B A: neither variant always runs on the cooler chip
    for terms, variants in order:
        x = x3 if terms == 3 else (x3[0], None)
        for name, flags in variants:
            for _ in range(2):
                trunk(x, terms, flags)
            torch.cuda.synchronize()
            clk = Clock()
            th = threading.Thread(target=clk.run, daemon=True)
            th.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                trunk(x, terms, flags)
            e1.record()
            torch.cuda.synchronize()
            clk.stop = True
            th.join()
            ms = e0.elapsed_time(e1) / reps
            print("terms=%d %-30s n=%d: %.3f ms per 32-layer trunk pass (%.1f us/layer), %.0f TFLOP/s algorithmic, "
                  "SM clock median %d MHz, power median %.0f W" % (terms, name, n, ms, ms / 32 * 1e3, flop / ms / 1e9,
                                                                  int(np.median(clk.v)), float(np.median(clk.p))),
                  flush=True)


if __name__ == "__main__":
    main()
