"""TF-V2 checkpoint reader/writer (dsin_b200/tf_checkpoint.py) -- host logic, no GPU.
The format restatement is pinned only against published known-answer values (CRC-32C check values of the
LevelDB/TensorFlow crc32c tests, the table magic) and hand-assembled byte streams; there is no TF checkpoint
in the reference tree to read (parity unpinned, see the module header)."""
import os
import struct

import numpy as np
import pytest

from dsin_b200 import synth
from dsin_b200 import tf_checkpoint as T


def test_crc32c_known_answers_and_chunked_path():
    assert T.crc32c(b"123456789") == 0xE3069283            # the standard CRC-32C check value
    assert T.crc32c(bytes(32)) == 0x8A9136AA               # rfc3720 B.4 vectors, also in leveldb's crc32c_test
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, size=300_007, dtype=np.uint8).tobytes()
    slow = T._raw_update(0xFFFFFFFF, big) ^ 0xFFFFFFFF      # byte-at-a-time definition
    assert T.crc32c(big) == slow                             # numpy chunk-parallel path + stitching
    assert T.crc32c(big[1234:], T.crc32c(big[:1234])) == slow  # continuation
    assert T.crc_unmask(T.crc_mask(slow)) == slow and T.crc_mask(slow) != slow
    assert T.crc_mask(0) == 0xA282EAD8


def test_snappy_hand_assembled_streams():
    # literal "abcd", copy-1 (len 4, offset 4), copy-2 (len 5, offset 8), overlapping copy-1 (len 7, offset 1)
    stream = bytes([20]) + bytes([3 << 2]) + b"abcd" + bytes([0b000_000_01, 4]) + \
        bytes([(5 - 1) << 2 | 2, 8, 0]) + bytes([(7 - 4) << 2 | 1, 1])
    assert T.snappy_uncompress(stream) == b"abcdabcd" + b"abcda" + b"a" * 7
    # long literal with an explicit one-byte length (tag 60)
    lit = bytes(range(100))
    assert T.snappy_uncompress(bytes([100]) + bytes([60 << 2, 99]) + lit) == lit
    with pytest.raises(T.CheckpointError):
        T.snappy_uncompress(bytes([4]) + bytes([0b000_000_01, 9]))  # copy before any output
    with pytest.raises(T.CheckpointError):
        T.snappy_uncompress(bytes([9]) + bytes([3 << 2]) + b"abcd")  # declared length mismatch


def test_table_roundtrip_multiblock_prefix_compression_and_snappy_block():
    items = [(b"", b"hdr")] + [(("scope/layer_%03d/weights" % i).encode(), os.urandom(7 + i % 5)) for i in range(200)]
    blob = T.write_table(items, block_size=256)   # forces many data blocks and restart points
    assert T.read_table(blob) == items
    assert struct.unpack("<Q", blob[-8:])[0] == 0xDB4775248B80FB57
    one = T.write_table(items[:3])
    assert T.read_table(one) == items[:3]
    assert T.read_table(T.write_table([])) == []
    with pytest.raises(T.CheckpointError):
        T.write_table([(b"b", b""), (b"a", b"")])
    # flip one payload byte -> block checksum must catch it
    bad = bytearray(blob)
    bad[10] ^= 1
    with pytest.raises(T.CheckpointError):
        T.read_table(bytes(bad))
    with pytest.raises(T.CheckpointError):
        T.read_table(blob[:-1] + b"\x00")
    # a table whose single data block is stored Snappy-compressed (type 1), assembled by hand
    b = T._BlockBuilder()
    b.add(b"k1", b"v1")
    b.add(b"k2", b"v2")
    raw = b.finish()
    comp = T._put_varint(len(raw)) + bytes([(len(raw) - 1) << 2]) + raw      # one literal element
    out = bytearray(comp) + bytes([1]) + struct.pack("<I", T.crc_mask(T.crc32c(comp + bytes([1]))))
    handle = T._put_varint(0) + T._put_varint(len(comp))

    def emit(block):
        h = T._put_varint(len(out)) + T._put_varint(len(block))
        out.extend(block + bytes([0]) + struct.pack("<I", T.crc_mask(T.crc32c(block + bytes([0])))))
        return h
    idx = T._BlockBuilder()
    idx.add(b"k2", handle)
    mh = emit(T._BlockBuilder().finish())
    ih = emit(idx.finish())
    out.extend(mh + ih + bytes(40 - len(mh + ih)) + struct.pack("<Q", T.TABLE_MAGIC))
    assert T.read_table(bytes(out)) == [(b"k1", b"v1"), (b"k2", b"v2")]


def test_bundle_entry_wire_format_is_the_published_proto():
    # BundleEntryProto{dtype=DT_FLOAT(1), shape{dim{size:3} dim{size:4}}, offset=48, size=48, crc32c=fixed32}
    enc = T._encode_entry(1, (3, 4), 0, 48, 48, 0x01020304)
    assert enc == bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x04,
                         0x20, 0x30, 0x28, 0x30, 0x35, 0x04, 0x03, 0x02, 0x01])
    e = T._parse_entry(enc)
    assert (e["dtype"], e["shape"], e["shard_id"], e["offset"], e["size"], e["crc32c"]) == (1, (3, 4), 0, 48, 48, 0x01020304)
    # BundleHeaderProto{num_shards=1, version{producer=1}}
    assert T._encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    assert T._parse_header(T._encode_header(1)) == {"num_shards": 1, "endianness": 0, "version": 1}


def test_checkpoint_roundtrip_of_the_model_variables(tmp_path):
    W = synth.make_weights(5)
    # what a training checkpoint additionally holds: optimizer slots, step counters (ignored on restore)
    extra = {"training-step/global_step": np.int64(1234), "beta1_power": np.float32(0.5),
             synth.ENC + "h1/weights/Adam": np.zeros((5, 5, 3, 64), np.float32)}
    prefix = str(tmp_path / "weights" / "model")
    T.write_checkpoint(prefix, {**W, **extra})
    assert sorted(os.listdir(tmp_path / "weights")) == ["checkpoint", "model.data-00000-of-00001", "model.index"]
    listing = T.list_variables(prefix)
    assert len(listing) == len(W) + 3 and listing[synth.ENC + "h1/weights"] == (np.dtype("float32"), (5, 5, 3, 64))
    names = synth.variable_names(5)
    assert names == sorted(W)
    R = T.read_checkpoint(prefix, names=names)
    assert sorted(R) == names
    for k in names:
        assert R[k].dtype == W[k].dtype and R[k].shape == W[k].shape and np.array_equal(R[k], W[k]), k
    everything = T.read_checkpoint(prefix)
    assert int(everything["training-step/global_step"]) == 1234 and everything["training-step/global_step"].shape == ()
    with pytest.raises(KeyError):
        T.read_checkpoint(prefix, names=["siNetwork/g_conv1/weights", "not/in/checkpoint"])
    with pytest.raises(FileNotFoundError):
        T.read_checkpoint(str(tmp_path / "nope"))
    # corrupt one tensor byte in the data shard -> the per-tensor CRC must reject it
    data = tmp_path / "weights" / "model.data-00000-of-00001"
    blob = bytearray(data.read_bytes())
    blob[len(blob) // 2] ^= 0x40
    data.write_bytes(bytes(blob))
    with pytest.raises(T.CheckpointError):
        T.read_checkpoint(prefix, names=names)
    # truncated data shard
    data.write_bytes(bytes(blob[:1000]))
    with pytest.raises(T.CheckpointError):
        T.read_checkpoint(prefix, names=names)
