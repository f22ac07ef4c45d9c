"""The CPU oracle of the PC1 entropy coder (oracle/pc_codec.c) checked against itself and against
the torch restatement of the probability model -- no GPU.  There is no reference bitstream to pin (the reference
ships helpers only, src/probclass_imgcomp.py:361-482): parity unpinned."""
import numpy as np
import pytest
import torch

from dsin_b200 import synth
from oracle import dsin_oracle as O
from oracle import pc_codec as P


@pytest.fixture(scope="module")
def weights():
    return synth.make_weights(0)


def test_exp_and_frequency_tables():
    for x in (0.0, -1e-3, -0.5, -1.0, -7.25, -30.0, -79.0):
        assert abs(P.exp_det(x) / np.exp(np.float64(x)) - 1) < 6e-6
    assert P.exp_det(-1000.0) == P.exp_det(-80.0) > 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        l = np.maximum(rng.normal(0, 4, 6), 0).astype(np.float32)  # ReLU'd logits like the model's
        f = P.freqs(l)
        assert f.sum() == 65536 and f.min() >= 1
        p = np.exp(l.astype(np.float64) - l.max())
        p /= p.sum()
        assert np.abs(f / 65536.0 - p).max() < 8.0 / 65536  # +1 floor per symbol, remainder (<= 6) to the arg-max
    assert P.freqs(np.zeros(6, np.float32)).tolist() == [10926, 10922, 10922, 10922, 10922, 10922]  # remainder -> first max
    assert P.freqs(np.array([0, 50, 0, 0, 0, 0], np.float32)).tolist() == [1, 65531, 1, 1, 1, 1]


def test_range_coder_carries_and_skewed_tables():
    rng = np.random.default_rng(1)
    n, L = 20000, 6
    tabs = np.ones((n, L), np.uint32)
    hot = rng.integers(0, L, n)
    tabs[np.arange(n), hot] = 65536 - (L - 1)            # extremely skewed tables
    sym = np.where(rng.random(n) < 0.97, hot, rng.integers(0, L, n)).astype(np.int32)
    bad, stream = P.rc_selftest(tabs, sym)
    assert bad == 0
    ideal = -np.log2(tabs[np.arange(n), sym] / 65536.0).sum()
    assert ideal - 32 <= 8 * len(stream) <= ideal + 16
    # tables that keep `low` near a byte boundary: long 0xFF runs and carry propagation
    tabs2 = np.tile(np.array([[1, 65534, 1]], np.uint32), (5000, 1))
    for pattern in ([2] * 5000, [1] * 4999 + [2], ([2] * 40 + [0]) * 121 + [1] * 39):
        bad, stream = P.rc_selftest(tabs2, np.array(pattern[:5000], np.int32))
        assert bad == 0
    bad, stream = P.rc_selftest(np.zeros((0, 3), np.uint32), np.zeros(0, np.int32))
    assert bad == 0 and stream == b""


@pytest.mark.parametrize("shape,nstreams", [((1, 1, 1), 1), ((2, 1, 3), 4), ((5, 4, 7), 3), ((32, 3, 2), 8), ((9, 6, 11), 8)])
def test_roundtrip_random_symbols(weights, shape, nstreams):
    rng = np.random.default_rng(7)
    sym = rng.integers(0, 6, size=shape).astype(np.int32)
    streams, ideal = P.encode(sym, weights, nstreams=nstreams)
    assert len(streams) == nstreams
    assert np.array_equal(P.decode(streams, shape, weights), sym)
    total = 8 * sum(len(s) for s in streams)
    assert ideal - 32 * nstreams <= total <= ideal + 16 * nstreams + 8   # trailing zero bytes are not stored
    # streams of depth slices that do not exist stay empty
    for k in range(shape[0], nstreams):
        assert streams[k] == b""


def test_code_length_matches_model_cross_entropy_and_stream_split(weights):
    x, _ = synth.make_batch(1, 80, 144, seed=3)
    enc = O.encode(torch.as_tensor(x), weights)
    sym = enc.symbols[0].numpy().astype(np.int32)
    ce_bits = float(O.probclass_bitcost(enc.qbar, enc.symbols, weights).sum())   # torch fp32 restatement
    one, ideal1 = P.encode(sym, weights, nstreams=1)
    eight, ideal8 = P.encode(sym, weights, nstreams=8)
    assert ideal1 == ideal8                                     # the model does not depend on the stream split
    assert abs(ideal1 - ce_bits) / ce_bits < 2e-4               # frequency quantisation only
    assert -32 <= 8 * len(one[0]) - ideal1 <= 16                # one stream: at most two bytes of overhead
    assert -32 * 8 <= 8 * sum(map(len, eight)) - ideal8 <= 16 * 8
    assert np.array_equal(P.decode(one, sym.shape, weights), sym)
    assert np.array_equal(P.decode(eight, sym.shape, weights), sym)
    # a damaged stream decodes to different symbols from the damaged point of that stream on, never crashes
    bad = list(eight)
    bad[2] = bytes([bad[2][0] ^ 0x55]) + bad[2][1:]
    dec = P.decode(bad, sym.shape, weights)
    assert not np.array_equal(dec, sym) and np.array_equal(dec[:2], sym[:2])
    assert dec.min() >= 0 and dec.max() < 6
