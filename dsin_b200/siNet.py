"""SI-Net (/root/reference/src/siNet.py:29-41): 9 dilated 3x3 convs (32 ch, LeakyReLU 0.2, bias,
no normalisation) + 1x1 conv to 3 channels, on libdsin_b200 kernels.

``siNet`` is a callable object so that it can own its weights (the reference's TF variables
live in the graph): ``siNet(input[N,6,H,W]) -> [N,3,H,W]``.  ``siNet.fused(x_dec_nhwc,
y_syn_nhwc)`` additionally fuses the normalise+concat in front (src/AE.py:67-68) and the
de-normalisation behind (src/AE.py:69)."""
from __future__ import annotations

import os

import numpy as np

from . import ops, synth


# "tc3": tcgen05 split-fp16 (fp32-class); "tc1": tcgen05 fp16; "simt": CUDA-core fp32
MODE = os.environ.get("DSIN_SINET_MODE", "tc3")
PAIR = os.environ.get("DSIN_SINET_PAIR", "1") != "0"  # pixel-pair form (128-byte TMA rows)
PAIR_SHARED = os.environ.get("DSIN_SINET_PAIR_SHARED", "1") != "0"  # even dilation: no MMAs on zero blocks


class SiNet(object):
    RATES = (1, 2, 4, 8, 16, 32, 64, 128, 1)

    def __init__(self):
        self.layers = None
        self._tc = None
        self.device = "cuda"

    def clone(self):
        """A fresh instance: every AE owns its SI-Net variables (as every TF graph of the reference does), so
        loading weights into one AE never invalidates what another AE's captured CUDA graphs point at."""
        return type(self)()

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
        self._tc = None
        # first layer with its 6 input channels zero-padded to 32 (tensor-core form)
        w1 = np.zeros((3, 3, 32, 32), dtype=np.float32)
        w1[:, :, :6, :] = W[S + "g_conv1/weights"]
        self._first_padded = ops.ConvLayer(w1, None, W[S + "g_conv1/biases"], dilation=1, act=ops.ACT_LRELU02,
                                           device=self.device)
        self._tc_first = None
        # even dilations: "pixel pair" form.  The NHWC tensor (n,H,W,32) is viewed as (n,H,W/2,64); a tap at
        # x offset +-d becomes +-d/2 pairs, and the 32x32 weight slab becomes a block-diagonal 64x64 one
        # (pixel parity is preserved by an even shift).  Twice the MMA work (on zeros) but half the TMA rows,
        # which is what bounds these layers.
        self._pair = {}
        for i, rate in enumerate(self.RATES):
            sc = S + "g_conv%d" % (i + 1)
            w = w1 if i == 0 else np.asarray(W[sc + "/weights"], np.float32)  # layer 0: cin padded to 32
            b = np.asarray(W[sc + "/biases"], np.float32)
            self._pair[i] = self._pair_layer(w, b, rate)
        self._pair_tc = {}

    def _pair_layer(self, w, b, rate):
        """3x3 (32->32, dilation `rate`) conv re-expressed on pixel pairs: out parity p at pair j reads input
        pixel 2j + p + dx, i.e. pair j + floor((p+dx)/2) with parity (p+dx) mod 2.  Even rates keep the parity
        (block-diagonal weights, pair taps at +-rate/2); odd rates mix parities (pair taps at +-(rate+1)/2 ...)."""
        offs = sorted({(p_ + (kx - 1) * rate) // 2 for kx in range(3) for p_ in (0, 1)})
        step = offs[1] - offs[0] if len(offs) > 1 else 1
        assert len(offs) == 3 and offs[2] - offs[1] == step and offs[1] == 0, offs  # symmetric 3-tap pattern
        wp = np.zeros((3, 3, 64, 64), dtype=np.float32)
        for kx in range(3):
            for p_ in (0, 1):
                src = p_ + (kx - 1) * rate
                po, pi = src // 2, src % 2
                wp[:, offs.index(po), pi * 32:(pi + 1) * 32, p_ * 32:(p_ + 1) * 32] += w[:, kx]
        layer = ops.ConvLayer(wp, None, np.concatenate([b, b]), dilation=rate, act=ops.ACT_LRELU02,
                              device=self.device)
        layer.dilation_x = step
        return layer

    def _run(self, net, post):
        n, hh, ww, _ = net.shape
        if MODE in ("tc3", "tc1") and hh >= 8 and ww >= 16:
            terms = 3 if MODE == "tc3" else 1
            if self._tc is None:
                self._tc = [ops.ConvTC(layer) for layer in self.layers[1:]]
            cur = ops.f32_to_split(ops.conv2d(net, self.layers[0]))  # g_conv1 has cin = 6: CUDA cores
            for tcl in self._tc[:-1]:
                cur = ops.conv_tc(cur, tcl, terms=terms)
            return ops.conv_tc(cur, self._tc[-1], terms=terms, out_f32=True, post=post)
        for layer in self.layers[:-1]:
            net = ops.conv2d(net, layer)
        return ops.conv2d(net, self.layers[-1], post=post)

    def __call__(self, input):  # noqa: A002 - reference argument name
        net = ops.nchw_to_nhwc(input.contiguous())
        return ops.nhwc_to_nchw(self._run(net, ops.POST_NONE))

    def fused(self, x_dec_nhwc, y_syn_nhwc):
        n, hh, ww, _ = x_dec_nhwc.shape
        if MODE in ("tc3", "tc1") and hh >= 8 and ww >= 16:
            terms = 3 if MODE == "tc3" else 1
            if self._tc is None:
                self._tc = [ops.ConvTC(layer) for layer in self.layers[1:]]
            if self._tc_first is None:
                self._tc_first = ops.ConvTC(self._first_padded)
            cur = ops.concat_normalize_split32(x_dec_nhwc, y_syn_nhwc)
            use_pair = PAIR and ww % 2 == 0 and ww // 2 >= 16
            for li, tcl in enumerate([self._tc_first] + self._tc[:-1]):
                if use_pair and li in self._pair:
                    if li not in self._pair_tc:
                        rate = self.RATES[li]
                        if PAIR_SHARED and li >= 1 and rate % 2 == 0:  # parity-preserving: shared 32x32 slab
                            self._pair_tc[li] = ops.PairSharedTC(self._tc[li - 1], rate)
                        else:
                            self._pair_tc[li] = ops.ConvTC(self._pair[li])
                    v = (cur[0].view(n, hh, ww // 2, 64), cur[1].view(n, hh, ww // 2, 64))
                    o = ops.conv_tc(v, self._pair_tc[li], terms=terms,
                                    prof=("tc%d_conv3x3_32to32_pair", 2.0 * n * hh * ww * 9 * (6 if li == 0 else 32) * 32))
                    cur = (o[0].view(n, hh, ww, 32), o[1].view(n, hh, ww, 32))
                else:
                    cur = ops.conv_tc(cur, tcl, terms=terms)
            out_nhwc = ops.conv_tc(cur, self._tc[-1], terms=terms, out_f32=True, post=ops.POST_DENORM)
            out = ops.nhwc_to_nchw(out_nhwc)
            out._dsin_nhwc = out_nhwc
            return out
        net = ops.concat_normalize(x_dec_nhwc, y_syn_nhwc)
        out_nhwc = self._run(net, ops.POST_DENORM)
        out = ops.nhwc_to_nchw(out_nhwc)
        out._dsin_nhwc = out_nhwc
        return out


siNet = SiNet()
