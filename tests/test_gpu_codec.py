"""PC1 entropy coder on the GPU (csrc/pc_codec.cu, through the C ABI) against the C oracle (oracle/pc_codec.c):
identical bytes, and each side decodes the other's streams.  Byte / integer work: bit-exact."""
import numpy as np
import pytest
import torch

from dsin_b200 import bitstream, synth
from oracle import dsin_oracle as O
from oracle import pc_codec as P

from parity_utils import calibrated_weights, make_ae

pytestmark = pytest.mark.gpu


def _model(weights):
    ae = make_ae(80, 144, weights)
    centers = torch.from_numpy(np.ascontiguousarray(weights[O.ENC + "centers"], np.float32)).cuda()
    return ae.pc_imgcomp, centers


def _split(blob):
    return bitstream.unpack(blob)[4]


@pytest.mark.parametrize("shape,nstreams", [((1, 1, 1), 1), ((5, 4, 7), 3), ((32, 10, 18), 8), ((9, 6, 159), 4)])
def test_bytes_equal_oracle_random_symbols(shape, nstreams):
    W = synth.make_weights(2)
    pc, centers = _model(W)
    rng = np.random.default_rng(5)
    sym = rng.integers(0, 6, size=(2,) + shape)
    blobs = pc.encode_symbols(torch.from_numpy(sym).cuda(), centers, nstreams=nstreams)
    for i in range(2):
        ref, _ideal = P.encode(sym[i].astype(np.int32), W, nstreams=nstreams)
        assert _split(blobs[i]) == ref, "GPU bitstream differs from the oracle's (image %d)" % i
    back = pc.decode_symbols(blobs, centers).cpu().numpy()
    assert np.array_equal(back, sym)


def test_real_symbols_cross_decode_and_code_length():
    """Symbols of a real encoder pass: GPU bytes == oracle bytes, GPU decodes oracle streams and vice versa, and
    the measured length sits on the cross-entropy estimate that `bitcost` reports (src/bits_imgcomp.py:13)."""
    W = calibrated_weights(0)
    pc, centers = _model(W)
    x, _ = synth.make_batch(2, 80, 144, seed=21)
    enc = O.encode(torch.as_tensor(x), W)
    sym = enc.symbols.numpy()
    blobs = pc.encode_symbols(enc.symbols.cuda(), centers)
    est_bits = O.probclass_bitcost(enc.qbar, enc.symbols, W).sum(dim=(1, 2, 3)).numpy()
    for i in range(2):
        ref, ideal = P.encode(sym[i].astype(np.int32), W)
        assert _split(blobs[i]) == ref
        real = bitstream.payload_bits(blobs[i])
        assert -32 * 8 <= real - ideal <= 16 * 8   # per stream: <= 2 bytes of flush, trailing zero bytes dropped
        assert abs(real - est_bits[i]) / est_bits[i] < 0.02, (real, est_bits[i])
        assert np.array_equal(P.decode(_split(blobs[i]), sym[i].shape, W), sym[i])   # oracle reads GPU streams
    ref_blobs = [bitstream.pack(P.encode(sym[i].astype(np.int32), W)[0], *sym[i].shape, 6) for i in range(2)]
    assert np.array_equal(pc.decode_symbols(ref_blobs, centers).cpu().numpy(), sym)  # GPU reads oracle streams


def test_full_size_batch_roundtrip_and_one_image_against_oracle():
    """BASELINE geometry (32 x 40 x 153 symbols per image), batch 8: encode -> decode is the identity
    (size-independent property); one image is also compared byte for byte with the oracle."""
    W = calibrated_weights(0)
    pc, centers = _model(W)
    x, _ = synth.make_batch(8, 320, 1224, seed=77)
    ae = make_ae(320, 1224, W)
    out = ae.reconstruct_device(torch.tensor(x).cuda(), torch.tensor(x).cuda())
    sym = out["symbols"]
    blobs = pc.encode_symbols(sym, centers)
    back = pc.decode_symbols(blobs, centers)
    assert torch.equal(back, sym)
    bits = np.array([bitstream.payload_bits(b) for b in blobs], np.float64)
    est = out["bits_sum"].cpu().numpy()
    assert np.all(np.abs(bits - est) / est < 0.005), (bits, est)
    ref, _ = P.encode(sym[3].cpu().numpy().astype(np.int32), W)
    assert _split(blobs[3]) == ref


def test_container_and_error_paths():
    W = synth.make_weights(2)
    pc, centers = _model(W)
    sym = torch.zeros((1, 4, 3, 5), dtype=torch.int64, device="cuda")
    blob = pc.encode_symbols(sym, centers, nstreams=2)[0]
    c, h, w, L, streams = bitstream.unpack(blob)
    assert (c, h, w, L, len(streams)) == (4, 3, 5, 6, 2)
    with pytest.raises(ValueError):
        bitstream.unpack(blob[:-1])
    with pytest.raises(ValueError):
        bitstream.unpack(b"XXXX" + blob[4:])
    with pytest.raises(ValueError):
        pc.decode_symbols([blob, pc.encode_symbols(torch.zeros((1, 4, 3, 6), dtype=torch.int64, device="cuda"), centers,
                                                   nstreams=2)[0]], centers)
    from dsin_b200 import ops
    with pytest.raises(RuntimeError, match="wider than 159"):   # wider than one CTA's step
        ops.pc_encode(torch.zeros((1, 2, 2, 160), dtype=torch.int64, device="cuda"), centers, pc._codec, 2)


def test_compress_decompress_through_the_facade():
    """Sender: AE.compress(x) -> bytes.  Receiver: AE.decompress(bytes, y) must reproduce what the one-call
    path computes from x and y: identical symbols, hence identical SI-Finder matches; images equal up to the
    qbar-vs-qhard rounding of the decoder input (documented in AE.decompress)."""
    W = calibrated_weights(0)
    ae = make_ae(80, 144, W)
    x, y = synth.make_batch(2, 80, 144, seed=8)
    x8, y8 = x.astype(np.uint8), y.astype(np.uint8)
    y_dec, y_syn, x_dec, x_with_si, bpp = [np.array(a) for a in ae.siNet_get_reconstructed(x8, y8)]
    sym_ref = ae.last["symbols"].clone()
    blobs = ae.compress(x8)
    assert all(isinstance(b, bytes) and b[:4] == b"DSPC" for b in blobs)
    real_bpp = 8.0 * sum(bitstream.payload_bits(b) // 8 for b in blobs) / (2 * 80 * 144)
    assert abs(real_bpp - float(bpp)) / float(bpp) < 0.03, (real_bpp, bpp)
    r_y_dec, r_y_syn, r_x_dec, r_x_with_si = ae.decompress(blobs, y8)
    assert torch.equal(ae.last["symbols"], sym_ref)
    assert np.array_equal(r_y_dec, y_dec)                       # the side image path does not involve the bitstream
    assert np.abs(r_x_dec - x_dec).max() < 2e-2                # qhard vs qbar: one fp32 rounding at the decoder input
    assert np.abs(r_x_with_si - x_with_si).max() < 0.5          # (0..255 scale; patch matches may move on near-ties)
    with pytest.raises(ValueError):
        ae.decompress(blobs[:1], y8)
