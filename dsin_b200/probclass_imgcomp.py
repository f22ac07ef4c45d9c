"""3-D masked-conv probability model (arch 'res_shallow') on libdsin_b200.

Mirrors /root/reference/src/probclass_imgcomp.py: ``get_network_cls``, ``_Network3D.bitcost``
(:63-106), ``auto_pad_value`` (:59-61), masks (:150-176).  The arithmetic-coding helpers (:361-482) have no
caller in the reference; ``encode_symbols`` / ``decode_symbols`` below are the coder they were written for
(PC1 format, csrc/pc_codec.cu, specified byte for byte in oracle/pc_codec.c).
"""
from __future__ import annotations

import numpy as np
import torch

from . import bitstream, ops, synth
from .Distortions_imgcomp import regularization_loss


# The two 24->24 layers and the head run on tcgen05; the all-CUDA-core kernel (ops.probclass_bits) is the
# cross-check the tests call directly.


def get_network_cls(pc_config):
    return {"res_shallow": _ResShallow}[pc_config.arch]


def create_masks(K=3):
    """first_mask / other_mask, DHW (src/probclass_imgcomp.py:150-176)."""
    first = np.ones((K // 2 + 1, K, K), dtype=np.float32)
    first[-1, K // 2, K // 2:] = 0
    first[-1, K // 2 + 1:, :] = 0
    other = np.ones((K // 2 + 1, K, K), dtype=np.float32)
    other[-1, K // 2, K // 2 + 1:] = 0
    other[-1, K // 2 + 1:, :] = 0
    return first, other


class _ResShallow(object):
    _PROBCLASS_SCOPE = "probclass3d"

    def __init__(self, pc_config, num_centers):
        self.config = pc_config
        self.L = num_centers
        if pc_config.kernel_size != 3:
            raise ValueError("only kernel_size 3 is built")
        self.first_mask, self.other_mask = create_masks(pc_config.kernel_size)
        self.weights = None
        self._tc = None
        self.device = "cuda"

    @classmethod
    def get_num_layers(cls):
        return 4

    @classmethod
    def get_context_size(cls, config):
        return cls.get_num_layers() * (config.kernel_size - 1) + 1

    def auto_pad_value(self, ae):
        if not self.config.use_centers_for_padding:
            return 0.0
        return float(ae.centers_host[0])

    def load_weights(self, W):
        P = synth.PC
        self._variables = W
        out = []
        for name, mask in (("conv3d_conv0_mask", self.first_mask), ("res1/conv3d_conv1_mask", self.other_mask),
                           ("res1/conv3d_conv2_mask", self.other_mask), ("conv3d_conv2_mask", self.other_mask)):
            w = W[P + name + "/weights"].astype(np.float32) * mask[..., None, None]
            b = W[P + name + "/biases"].astype(np.float32)
            out.append((torch.from_numpy(np.ascontiguousarray(w)).to(self.device),
                        torch.from_numpy(np.ascontiguousarray(b)).to(self.device)))
        self.weights = out
        self._tc = None
        # live-tap-major fp32 slices for the entropy coder: (taps, cin, cout), taps in (kd, kh, kw) raster order
        self._codec = []
        for i, name in enumerate(("conv3d_conv0_mask", "res1/conv3d_conv1_mask", "res1/conv3d_conv2_mask",
                                  "conv3d_conv2_mask")):
            w = np.asarray(W[P + name + "/weights"], np.float32)
            taps = [(0, kh, kw) for kh in range(3) for kw in range(3)] + [(1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 1, 0)]
            if i > 0:
                taps.append((1, 1, 1))  # every layer but the first also sees the current position
            packed = np.ascontiguousarray(np.stack([w[kd, kh, kw] for kd, kh, kw in taps]))
            self._codec.append(torch.from_numpy(packed).to(self.device))
            self._codec.append(torch.from_numpy(np.ascontiguousarray(W[P + name + "/biases"], np.float32)).to(self.device))

    def regularization_loss(self):
        """src/probclass_imgcomp.py:88-95,115-119: None unless the config sets a regularization_factor; then
        factor * sum of l2_loss over the conv3d weights that `tf.losses.get_regularization_loss(scope='probclass3d')`
        selects (prefix rule, see Distortions_imgcomp.regularization_loss)."""
        if self.config.regularization_factor is None:
            return None
        return regularization_loss(self._variables, self._PROBCLASS_SCOPE, self.config.regularization_factor)

    def bitcost(self, q, target_symbols, is_training=False, pad_value=0, terms=3):
        """q: qbar NCHW fp32, target_symbols NCHW int64 -> bits per symbol NCHW.  The fp64
        per-image sums ride along as ``._dsin_sum`` for bits_imgcomp.bitcost_to_bpp."""
        if is_training is True:
            raise NotImplementedError("dsin_b200 implements the inference path only")
        if q.dim() != 4:
            raise ValueError("expected NCHW, got {}".format(tuple(q.shape)))
        n, c, hh, ww = q.shape
        if hh + 2 < 8 or ww + 2 < 16 or self.config.arch_param__k != 24:
            raise ValueError("dsin_b200 probclass: needs arch_param__k == 24 and a bottleneck of at least 6x14")
        if self._tc is None:
            self._tc = ops.ProbclassTC(self.weights)
        bits, sums = ops.probclass_bits_tc(q.contiguous(), target_symbols.contiguous(), self._tc,
                                           float(pad_value), terms=terms)
        bits._dsin_sum = sums
        return bits

    # ------------------------------------------------------------------ real entropy coding (PC1)
    def _check_codec(self):
        if not self.config.use_centers_for_padding:
            raise NotImplementedError("the PC1 coder pads with centres[0] (use_centers_for_padding = True)")
        if self.config.arch_param__k != 24:
            raise NotImplementedError("the PC1 coder is built for 24 hidden channels")

    def encode_symbols(self, symbols, centers, nstreams=8):
        """symbols (n,c,h,w) int64 CUDA, centers (L,) fp32 CUDA -> list of n self-describing bitstreams (bytes).
        Each symbol is range-coded with the frequencies the context model predicts from the symbols before it
        (src/probclass_imgcomp.py:421-470: get_freqs), so len(bitstream) is the real code length."""
        self._check_codec()
        n, c, hh, ww = symbols.shape
        out, sizes, status = ops.pc_encode(symbols.contiguous(), centers, self._codec, nstreams)
        sizes_h = sizes.cpu().numpy()
        if int(status.item()) != 0:
            raise RuntimeError("PC1 encoder: a stream exceeded its capacity")
        out_h = out.cpu().numpy()
        return [bitstream.pack([out_h[i, k, :sizes_h[i, k]].tobytes() for k in range(nstreams)], c, hh, ww, self.L)
                for i in range(n)]

    def decode_symbols(self, bitstreams, centers, expect_shape=None):
        """list of n bitstreams (same shape) -> symbols (n,c,h,w) int64 CUDA.  expect_shape = (c, h, w) the caller's
        model geometry implies; a header that disagrees is rejected before anything is allocated."""
        self._check_codec()
        parsed = [bitstream.unpack(b) for b in bitstreams]
        c, hh, ww, L, streams0 = parsed[0]
        nstreams = len(streams0)
        for pc, ph, pw, pL, st in parsed:
            if (pc, ph, pw, pL, len(st)) != (c, hh, ww, L, nstreams):
                raise ValueError("bitstreams of one batch must share their geometry")
        if L != self.L:
            raise ValueError("bitstream has {} centres, model has {}".format(L, self.L))
        # the header is untrusted input: refuse geometries the model cannot have produced BEFORE allocating for them
        if expect_shape is not None and (c, hh, ww) != tuple(expect_shape):
            raise ValueError("bitstream holds a {}x{}x{} symbol volume, this model / image size needs {}x{}x{}".format(
                c, hh, ww, *expect_shape))
        if not (1 <= c <= 256 and 1 <= hh <= 58 and 1 <= ww <= 159 and 1 <= nstreams <= 64):
            raise ValueError("bitstream geometry {}x{}x{} / {} streams is outside what the PC1 coder supports".format(
                c, hh, ww, nstreams))
        for p_ in parsed:
            for s_ in p_[4]:
                if len(s_) > 2 * ((c + nstreams - 1) // nstreams) * hh * ww + 16:
                    raise ValueError("a stream is longer than its symbols can need (16 bits per symbol)")
        cap = max(max(len(s) for s in p[4]) for p in parsed) + 8
        buf = np.zeros((len(parsed), nstreams, cap), np.uint8)
        sizes = np.zeros((len(parsed), nstreams), np.int64)
        for i, p in enumerate(parsed):
            for k, s in enumerate(p[4]):
                buf[i, k, :len(s)] = np.frombuffer(s, np.uint8)
                sizes[i, k] = len(s)
        dev = centers.device
        return ops.pc_decode(torch.from_numpy(buf).to(dev), torch.from_numpy(sizes).to(dev), (len(parsed), c, hh, ww),
                             centers, self._codec)
