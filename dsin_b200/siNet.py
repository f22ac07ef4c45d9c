"""SI-Net (/root/reference/src/siNet.py:29-41): 9 dilated 3x3 convs (32 ch, LeakyReLU 0.2, bias,
no normalisation) + 1x1 conv to 3 channels, on libdsin_b200 kernels.

``siNet`` is a callable object so that it can own its weights (the reference's TF variables
live in the graph): ``siNet(input[N,6,H,W]) -> [N,3,H,W]``.  ``siNet.fused(x_dec_nhwc,
y_syn_nhwc)`` additionally fuses the normalise+concat in front (src/AE.py:67-68) and the
de-normalisation behind (src/AE.py:69)."""
from __future__ import annotations

import numpy as np

from . import ops, synth


class SiNet(object):
    RATES = (1, 2, 4, 8, 16, 32, 64, 128, 1)

    def __init__(self):
        self.layers = None
        self.device = "cuda"

    def load_weights(self, W):
        S = synth.SIN
        layers = []
        for i, rate in enumerate(self.RATES):
            sc = S + "g_conv%d" % (i + 1)
            layers.append(ops.ConvLayer(W[sc + "/weights"], None, W[sc + "/biases"], dilation=rate,
                                        act=ops.ACT_LRELU02, device=self.device))
        sc = S + "g_conv_last"
        layers.append(ops.ConvLayer(W[sc + "/weights"], None, W[sc + "/biases"], act=ops.ACT_NONE,
                                    device=self.device))
        self.layers = layers

    def _run(self, net, post):
        for layer in self.layers[:-1]:
            net = ops.conv2d(net, layer)
        return ops.conv2d(net, self.layers[-1], post=post)

    def __call__(self, input):  # noqa: A002 - reference argument name
        net = ops.nchw_to_nhwc(input.contiguous())
        return ops.nhwc_to_nchw(self._run(net, ops.POST_NONE))

    def fused(self, x_dec_nhwc, y_syn_nhwc):
        net = ops.concat_normalize(x_dec_nhwc, y_syn_nhwc)
        out_nhwc = self._run(net, ops.POST_DENORM)
        out = ops.nhwc_to_nchw(out_nhwc)
        out._dsin_nhwc = out_nhwc
        return out


siNet = SiNet()
