"""Property-based tests (hypothesis) of the byte-level host / oracle code: range coder, bitstream container,
TF-V2 checkpoint tables.  No GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

from dsin_b200 import bitstream
from dsin_b200 import tf_checkpoint as T
from oracle import pc_codec as P


@st.composite
def _tables_and_symbols(draw):
    L = draw(st.integers(2, 8))
    n = draw(st.integers(0, 300))
    seed = draw(st.integers(0, 2 ** 32 - 1))
    rng = np.random.default_rng(seed)
    skew = draw(st.sampled_from([0.1, 1.0, 8.0]))
    w = rng.gamma(skew, size=(n, L)) + 1e-9
    f = np.maximum(1, np.floor(w / w.sum(1, keepdims=True) * (65536 - L)).astype(np.int64))
    f[np.arange(n), f.argmax(1)] += 65536 - f.sum(1)            # every table sums to 2^16 with entries >= 1
    sym = np.array([rng.choice(L, p=row / 65536.0) for row in f], dtype=np.int32) if n else np.zeros(0, np.int32)
    return f.astype(np.uint32), sym


@settings(max_examples=60, deadline=None)
@given(_tables_and_symbols())
def test_range_coder_roundtrip_and_length(ts):
    tables, sym = ts
    bad, stream = P.rc_selftest(tables, sym)
    assert bad == 0
    ideal = -np.log2(tables[np.arange(sym.size), sym] / 65536.0).sum() if sym.size else 0.0
    assert ideal - 32 <= 8 * len(stream) <= ideal + 16


@settings(max_examples=60, deadline=None)
@given(st.lists(st.binary(max_size=40), min_size=1, max_size=12), st.integers(1, 512), st.integers(1, 300),
       st.integers(1, 300), st.integers(2, 8))
def test_container_roundtrip(streams, c, h, w, L):
    blob = bitstream.pack(streams, c, h, w, L)
    assert bitstream.unpack(blob) == (c, h, w, L, streams)
    assert bitstream.payload_bits(blob) == 8 * sum(map(len, streams))


_names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_/0123456789", min_size=1, max_size=40)


@settings(max_examples=40, deadline=None)
@given(st.dictionaries(st.binary(min_size=1, max_size=60), st.binary(max_size=80), max_size=120), st.integers(64, 4096))
def test_sorted_table_roundtrip(items, block_size):
    ordered = sorted(items.items())
    assert T.read_table(T.write_table(ordered, block_size=block_size)) == ordered


@settings(max_examples=15, deadline=None)
@given(spec=st.dictionaries(_names, st.tuples(st.sampled_from(["<f4", "<f8", "<i4", "<i8", "u1", "<f2"]),
                                              st.lists(st.integers(0, 5), max_size=4)), min_size=1, max_size=12),
       seed=st.integers(0, 2 ** 31))
def test_checkpoint_roundtrip_random_variables(spec, seed, tmp_path_factory):
    rng = np.random.default_rng(seed)
    variables = {k: (rng.normal(size=shape) * 50).astype(np.dtype(dt)) for k, (dt, shape) in spec.items()}
    prefix = str(tmp_path_factory.mktemp("ck") / "model")
    T.write_checkpoint(prefix, variables)
    back = T.read_checkpoint(prefix)
    assert sorted(back) == sorted(variables)
    for k, v in variables.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k


@settings(max_examples=25, deadline=None)
@given(c=st.integers(1, 9), h=st.integers(1, 7), w=st.integers(1, 9), nstreams=st.integers(1, 8),
       seed=st.integers(0, 2 ** 31), scale=st.sampled_from([0.3, 1.0, 6.0]))
def test_oracle_codec_roundtrip_random_models(c, h, w, nstreams, seed, scale):
    """Any weights (including sharply peaked models, scale 6), any small geometry, any stream count: decode(encode)
    is the identity and the stream length tracks the model's own code length."""
    from dsin_b200 import synth
    rng = np.random.default_rng(seed)
    W = synth.make_weights(seed % 7)
    for k in list(W):
        if k.startswith(P.PC) and k.endswith("/weights"):
            W[k] = (W[k] * scale).astype(np.float32)
    sym = rng.integers(0, 6, size=(c, h, w)).astype(np.int32)
    streams, ideal = P.encode(sym, W, nstreams=nstreams)
    assert np.array_equal(P.decode(streams, sym.shape, W), sym)
    total = 8 * sum(len(s) for s in streams)
    assert ideal - 32 * nstreams <= total <= ideal + 16 * nstreams + 8


def _snappy_compress(data):
    """A tiny greedy raw-Snappy encoder (literals + copies with 1-, 2- and 4-byte offsets) to feed the decoder."""
    out = bytearray(T._put_varint(len(data)))
    lit = bytearray()

    def flush_literal():
        i = 0
        while i < len(lit):
            chunk = lit[i:i + 70000]
            n = len(chunk) - 1
            if n < 60:
                out.append(n << 2)
            elif n < 256:
                out.extend([60 << 2, n])
            elif n < 65536:
                out.extend([61 << 2]) or out.extend(n.to_bytes(2, "little"))
            else:
                out.extend([62 << 2]) or out.extend(n.to_bytes(3, "little"))
            out.extend(chunk)
            i += len(chunk)
        lit.clear()

    table, i = {}, 0
    while i < len(data):
        key = bytes(data[i:i + 4])
        j = table.get(key) if len(key) == 4 else None
        if len(key) == 4:
            table[key] = i
        if j is not None and i - j > 0:
            ln = 4
            while i + ln < len(data) and ln < 64 and data[j + ln] == data[i + ln]:  # may overlap: run-length style
                ln += 1
            off = i - j
            flush_literal()
            if 4 <= ln <= 11 and off < 2048:
                out.extend([((off >> 8) << 5) | ((ln - 4) << 2) | 1, off & 0xFF])
            elif off < 65536:
                out.extend([((ln - 1) << 2) | 2]) or out.extend(off.to_bytes(2, "little"))
            else:
                out.extend([((ln - 1) << 2) | 3]) or out.extend(off.to_bytes(4, "little"))
            i += ln
        else:
            lit.append(data[i])
            i += 1
    flush_literal()
    return bytes(out)


@settings(max_examples=80, deadline=None)
@given(st.lists(st.one_of(st.binary(max_size=30), st.integers(1, 300).map(lambda n: b"ab" * n),
                          st.integers(1, 400).map(lambda n: bytes([n % 251]) * n)), max_size=12))
def test_snappy_decoder_inverts_a_real_encoder(parts):
    data = b"".join(parts)
    comp = _snappy_compress(data)
    assert T.snappy_uncompress(comp) == data
