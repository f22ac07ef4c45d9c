"""Encoder/decoder networks (arch 'CVPR') on libdsin_b200 kernels.

Mirrors the interface of /root/reference/src/autoencoder_imgcomp.py: ``get_network_cls``,
``_Network.encode/decode/get_centers_variable``, ``EncoderOutput``; the bodies
(:219-269, residual_block :275-288, BN scopes :106-125, heatmap :173-201, quantiser via
quantizer_imgcomp.py:43-95) run as CUDA kernels on NHWC activations.  Inference only:
BatchNorm is folded to a per-channel scale/shift at weight-load time.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from . import ops, synth
from .Distortions_imgcomp import regularization_loss

EncoderOutput = namedtuple("EncoderOutput", ["qbar", "qhard", "symbols", "z", "heatmap"])

BN_EPS = np.float32(1e-5)
arch_param_n = 128
SCOPE_AE = "autoencoder"                 # src/autoencoder_imgcomp.py:21-23
SCOPE_AE_ENC = SCOPE_AE + "/encoder"
SCOPE_AE_DEC = SCOPE_AE + "/decoder"

# Every conv of the encoder/decoder runs on tcgen05 (csrc/conv_tc*.cu).  `terms` (precision.py) says how many MMAs
# build one product: 3 = split-fp16 hi*hi + hi*lo + lo*hi (fp32-class), 1 = fp16 operands.  There is no other
# backend: an image whose quarter-resolution trunk is smaller than one 16x32 tile block is rejected.


def get_network_cls(config):
    return {"CVPR": _CVPR}[config.arch]


def _fold_bn(W, scope):
    g = W[scope + "/BatchNorm/gamma"].astype(np.float32)
    b = W[scope + "/BatchNorm/beta"].astype(np.float32)
    m = W[scope + "/BatchNorm/moving_mean"].astype(np.float32)
    v = W[scope + "/BatchNorm/moving_variance"].astype(np.float32)
    s = (g / np.sqrt(v + BN_EPS)).astype(np.float32)
    return s, (b - m * s).astype(np.float32)


class _Network(object):
    def __init__(self, config, quantize=True):
        self.config = config
        self.quantize = quantize
        self.num_chan_bn_including_heatmap = config.num_chan_bn + 1
        self._centers = None
        self.layers = {}
        self._tc_layers = {}
        self._variables = {}
        self.device = "cuda"

    @staticmethod
    def get_subsampling_factor():
        raise NotImplementedError()

    def get_centers_variable(self):
        if self._centers is None:
            raise ValueError("Call load_weights(...) before trying to access centers")
        return self._centers

    def encoder_regularization_loss(self):
        """includes centers regularization (src/autoencoder_imgcomp.py:79-82); see Distortions_imgcomp for the
        scope rule that makes this 0.0 in the graph src/AE.py builds."""
        return regularization_loss(self._variables, SCOPE_AE_ENC, self.config.regularization_factor,
                                   self.config.regularization_factor_centers)

    def decoder_regularization_loss(self):
        return regularization_loss(self._variables, SCOPE_AE_DEC, self.config.regularization_factor)

    # -- weights ---------------------------------------------------------------------------
    def _conv(self, W, scope, stride=1, relu=True, transposed=False, post=ops.POST_NONE):
        w = W[scope + "/weights"]
        if transposed:  # reference layout [k,k,out,in] -> [k,k,in,out]
            w = np.transpose(w, (0, 1, 3, 2))
        s, t = _fold_bn(W, scope)
        self.layers[scope] = ops.ConvLayer(w, s, t, stride=stride, transposed=transposed,
                                           act=ops.ACT_RELU if relu else ops.ACT_NONE, post=post,
                                           device=self.device)

    def _transposed_as_conv3x3(self, layer):
        """A k=5 stride-2 TF-SAME transposed conv (cout = 3) as ONE 3x3 stride-1 conv to 12 phase-channels
        (2x2 sub-pixel phases x 3 colours) + depth-to-space: out[2a+py, 2b+px] reads in[a+dy, b+dx] with
        ky = py + 1 - 2dy, kx = px + 1 - 2dx, so every phase only touches the 3x3 neighbourhood of (a, b).
        The four-phase form fetches the activation tile 25 times, this one 9 times."""
        w = layer.w.cpu().numpy()  # [5,5,cin,3]
        k, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        assert k == 5 and cout == 3
        w9 = np.zeros((3, 3, cin, 12), dtype=np.float32)
        for py in range(2):
            for px in range(2):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        ky, kx = py + 1 - 2 * dy, px + 1 - 2 * dx
                        if 0 <= ky < k and 0 <= kx < k:
                            w9[dy + 1, dx + 1, :, (py * 2 + px) * 3:(py * 2 + px) * 3 + 3] = w[ky, kx]
        scale = np.tile(layer.scale.cpu().numpy(), 4)
        shift = np.tile(layer.shift.cpu().numpy(), 4)
        return ops.ConvLayer(w9, scale, shift, act=layer.act, post=ops.POST_DENORM_CLIP_D2S, device=self.device)

    def _stem_as_conv3x3(self, layer):
        """The 5x5 stride-2 TF-SAME stem (cin = 3) as a 3x3 stride-1 conv over the space-to-depth(2) image
        (12 channels, padded to 32): input pixel (2(a+dy)+sy, 2(b+dx)+sx) = (2a+ky-1, 2b+kx-1), i.e.
        ky = 2dy+sy+1, kx = 2dx+sx+1."""
        w = layer.w.cpu().numpy()  # [5,5,3,cout]
        k, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        assert k == 5 and cin == 3
        w9 = np.zeros((3, 3, 32, cout), dtype=np.float32)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                for sy in range(2):
                    for sx in range(2):
                        ky, kx = 2 * dy + sy + 1, 2 * dx + sx + 1
                        if 0 <= ky < k and 0 <= kx < k:
                            w9[dy + 1, dx + 1, (sy * 2 + sx) * 3:(sy * 2 + sx) * 3 + 3, :] = w[ky, kx]
        return ops.ConvLayer(w9, layer.scale.cpu().numpy(), layer.shift.cpu().numpy(), act=layer.act,
                             device=self.device)

    def load_weights(self, W):
        raise NotImplementedError()

    def encode(self, x, is_training=False, terms=3):
        if is_training is True:
            raise NotImplementedError("dsin_b200 implements the inference path only")
        return self._encode(x, terms)

    def decode(self, q, is_training=False, terms=3):
        if is_training is True:
            raise NotImplementedError("dsin_b200 implements the inference path only")
        return self._decode(q, terms)


class _CVPR(_Network):
    @staticmethod
    def get_subsampling_factor():
        return 8

    def load_weights(self, W):
        B = self.config.arch_param_B
        E, D = synth.ENC, synth.DEC
        self._variables = W
        self._tc_layers = {}
        self._h13_tc = None
        self._centers = torch.from_numpy(np.ascontiguousarray(W[E + "centers"], np.float32)).to(self.device)
        self.centers_host = np.asarray(W[E + "centers"], np.float32).copy()
        self._conv(W, E + "h1", stride=2)
        self._conv(W, E + "h2", stride=2)
        self._conv(W, E + "to_bn", stride=2, relu=False)
        self._conv(W, D + "from_bn", stride=2, transposed=True)
        self._conv(W, D + "h12", stride=2, transposed=True)
        self._conv(W, D + "h13", stride=2, transposed=True, relu=False, post=ops.POST_DENORM_CLIP)
        self._h13_d2s = self._transposed_as_conv3x3(self.layers[D + "h13"])
        self._h1_s2d = self._stem_as_conv3x3(self.layers[E + "h1"])
        self._h1_tc = None
        for pre, blk, fin in ((E, "res_block_enc_%d/enc_%d_%d", "res_block_enc_final"),
                              (D, "res_block_dec_%d/dec_%d_%d", "dec_after_res")):
            for b in range(B):
                for i in (1, 2, 3):
                    sc = pre + blk % (b, b, i)
                    self._conv(W, sc + "/conv1", relu=True)
                    self._conv(W, sc + "/conv2", relu=False)
            self._conv(W, pre + fin + "/conv1", relu=False)  # activation_fn=None on both
            self._conv(W, pre + fin + "/conv2", relu=False)

    # -- tensor-core path --------------------------------------------------------------------
    def _tc(self, scope):
        t = self._tc_layers.get(scope)
        if t is None:
            t = self._tc_layers[scope] = ops.ConvTC(self.layers[scope])
        return t

    @staticmethod
    def _require_tileable(hh, ww):
        """All tcgen05 layers tile 8x16 pixel blocks (16x32 input pixels for the stride-2 layers)."""
        if hh < 16 or ww < 32 or hh % 2 or ww % 2:
            raise ValueError("dsin_b200: images must be at least 64x128 with sides divisible by 8 "
                             "(quarter-resolution trunk %dx%d); there is no non-tensor-core path" % (hh, ww))

    def _trunk_tc(self, cur, pre, blk, fin, terms):
        """cur: split-fp16 pair.  15 residual blocks + final block, 3 skip levels."""
        r0 = cur
        for b in range(self.config.arch_param_B):
            rb = cur
            for i in (1, 2, 3):
                sc = pre + blk % (b, b, i)
                t = ops.conv_tc(cur, self._tc(sc + "/conv1"), terms=terms)
                cur = ops.conv_tc(t, self._tc(sc + "/conv2"), res1=cur, res2=rb if i == 3 else None, terms=terms)
        t = ops.conv_tc(cur, self._tc(pre + fin + "/conv1"), terms=terms)
        return ops.conv_tc(t, self._tc(pre + fin + "/conv2"), res1=cur, res2=r0, terms=terms)

    def _encode_tc(self, x, terms):
        E = synth.ENC
        if self._h1_tc is None:
            self._h1_tc = ops.ConvTC(self._h1_s2d)
        cur = ops.conv_tc(ops.nchw_to_s2d_split32(x, with_lo=terms == 3), self._h1_tc, terms=terms,
                          prof=("tc%d_conv5x5_3to64_s2_as3x3", 2.0 * x.shape[0] * (x.shape[2] // 2)
                                * (x.shape[3] // 2) * 25 * 3 * 64))
        cur = ops.conv_tc(cur, self._tc(E + "h2"), terms=terms)
        cur = self._trunk_tc(cur, E, "res_block_enc_%d/enc_%d_%d", "res_block_enc_final", terms)
        return ops.conv_tc(cur, self._tc(E + "to_bn"), terms=terms, out_f32=True)

    def _decode_tc(self, q_nhwc, terms):
        D = synth.DEC
        cur = ops.f32_to_split(q_nhwc, with_lo=terms == 3)
        cur = ops.conv_tc(cur, self._tc(D + "from_bn"), terms=terms)
        cur = self._trunk_tc(cur, D, "res_block_dec_%d/dec_%d_%d", "dec_after_res", terms)
        cur = ops.conv_tc(cur, self._tc(D + "h12"), terms=terms)
        if self._h13_tc is None:
            self._h13_tc = ops.ConvTC(self._h13_d2s)
        return ops.conv_tc(cur, self._h13_tc, terms=terms, out_f32=True,  # BN, denormalise, clip, depth-to-space
                           prof=("tc%d_conv5x5_64to3_T_as3x3", 2.0 * cur[0].shape[0] * cur[0].shape[1]
                                 * cur[0].shape[2] * 25 * 64 * 3))

    def _encode(self, x, terms=3):
        """x: (N,3,H,W) fp32 CUDA tensor, uint8-valued -> EncoderOutput(qbar, qhard, symbols, z, heatmap), all
        (N,C,h,w) like the reference's (src/autoencoder_imgcomp.py:239-245)."""
        self._require_tileable(x.shape[2] // 4, x.shape[3] // 4)
        z33 = self._encode_tc(x, terms)
        qbar_nhwc, qbar_nchw, symbols, qhard, z, heatmap = ops.heatmap_quantize(z33, self._centers, full=True)
        qbar_nchw._dsin_nhwc = qbar_nhwc
        return EncoderOutput(qbar_nchw, qhard, symbols, z, heatmap)

    def _decode(self, q, terms=3):
        """q: qbar (N,C,h,w); returns x_dec (N,3,H,W) clipped to [0,255]."""
        q_nhwc = getattr(q, "_dsin_nhwc", None)
        if q_nhwc is None:
            q_nhwc = ops.nchw_to_nhwc(q.contiguous(), normalize=False)
        self._require_tileable(2 * q.shape[2], 2 * q.shape[3])
        img_nhwc = self._decode_tc(q_nhwc, terms)
        out = ops.nhwc_to_nchw(img_nhwc)
        out._dsin_nhwc = img_nhwc
        return out
