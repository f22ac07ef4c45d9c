"""TensorFlow V2 checkpoint ("tensor bundle") reader / writer without TensorFlow.

The reference saves and restores its weights with `tf.train.Saver` (src/AE.py:154-175); the authors'
released models are TF-V2 checkpoints `weights/<name>/model.{index,data-00000-of-00001}`
(README.md:37-39).  TensorFlow 1.11 is not installable here, so this module restates the two published
on-disk formats involved:

* `<prefix>.index` is a LevelDB-style sorted string table (data blocks with prefix-compressed keys and
  restart arrays, an index block, a 48-byte footer ending in the magic 0xdb4775248b80fb57; every block is
  followed by a 1-byte compression type and a masked CRC-32C).  Key "" holds a `BundleHeaderProto`
  (num_shards, endianness, version); every other key is a variable name whose value is a
  `BundleEntryProto` (dtype, shape, shard_id, offset, size, masked crc32c of the tensor bytes).
* `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at those offsets.

Blocks may be Snappy-compressed (TensorFlow's table writer compresses when it saves >= 12.5 %), so a
Snappy decompressor is included.  Every CRC (block trailers and tensor payloads) is verified on read.

PARITY UNPINNED: there is no TF checkpoint under /root/reference and no TensorFlow to produce one; the
reader is tested against this module's own writer, hand-assembled Snappy streams and the published CRC-32C
check value.  Host-side logic only -- nothing here touches the GPU.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
NO_COMPRESSION, SNAPPY_COMPRESSION = 0, 1
RESTART_INTERVAL = 16
BLOCK_SIZE = 262144
CRC_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
          6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
          22: np.dtype("<u4"), 23: np.dtype("<u8")}
DTYPE_IDS = {v: k for k, v in DTYPES.items()}
DT_STRING = 7


class CheckpointError(ValueError):
    pass


# ----------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _make_crc_table():
    poly = 0x82F63B78
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t[i] = c
    return t


_CRC_T = _make_crc_table()
_CRC_TL = [int(v) for v in _CRC_T]
_SHIFT_CACHE = {}
_CHUNK = 4096


def _raw_update(state, data):
    """CRC register after feeding `data` (bytes) starting from `state`; no init / final xor."""
    t = _CRC_TL
    for b in data:
        state = t[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def _shift_tables(n):
    """Four 256-entry tables of the linear map  state -> register after n zero bytes."""
    tabs = _SHIFT_CACHE.get(n)
    if tabs is None:
        zeros = bytes(n)
        basis = [_raw_update(1 << b, zeros) for b in range(32)]
        tabs = []
        for byte in range(4):
            tab = [0] * 256
            for v in range(256):
                acc = 0
                for bit in range(8):
                    if v >> bit & 1:
                        acc ^= basis[8 * byte + bit]
                tab[v] = acc
            tabs.append(tab)
        _SHIFT_CACHE[n] = tabs
    return tabs


def crc32c(data, crc=0):
    """CRC-32C of `data` (bytes-like), continuing from a previous value `crc`.
    Large inputs are processed as many 4 KiB chunks in parallel with numpy and stitched together through
    the linearity of the CRC register (register(s, A) = register(0, A) xor shift_|A|(s))."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8)
    state = (crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
    nfull = buf.size // _CHUNK
    if nfull >= 8:
        rows = buf[:nfull * _CHUNK].reshape(nfull, _CHUNK)
        reg = np.zeros(nfull, dtype=np.uint32)
        for j in range(_CHUNK):
            reg = _CRC_T[(reg ^ rows[:, j]) & np.uint32(0xFF)] ^ (reg >> np.uint32(8))
        t0, t1, t2, t3 = _shift_tables(_CHUNK)
        for r in reg.tolist():
            state = (t0[state & 0xFF] ^ t1[(state >> 8) & 0xFF] ^ t2[(state >> 16) & 0xFF] ^ t3[state >> 24]) ^ r
        tail = buf[nfull * _CHUNK:]
    else:
        tail = buf
    state = _raw_update(state, tail.tobytes())
    return state ^ 0xFFFFFFFF


def crc_mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + CRC_MASK_DELTA) & 0xFFFFFFFF


def crc_unmask(masked):
    rot = (masked - CRC_MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varints / protobuf wire format
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yield (field_number, wire_type, value) of one protobuf message; value is int or bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            val, pos = bytes(buf[pos:pos + n]), pos + n
            if len(val) != n:
                raise CheckpointError("truncated protobuf field")
        elif wt == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf):
    dims = []
    for field, _wt, val in _proto_fields(buf):
        if field == 2:  # TensorShapeProto.Dim
            size = 0
            for f2, _w2, v2 in _proto_fields(val):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif field == 3 and val:
            raise CheckpointError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _wt, val in _proto_fields(buf):
        if field == 1:
            e["dtype"] = val
        elif field == 2:
            e["shape"] = _parse_shape(val)
        elif field == 3:
            e["shard_id"] = val
        elif field == 4:
            e["offset"] = _signed64(val)
        elif field == 5:
            e["size"] = _signed64(val)
        elif field == 6:
            e["crc32c"] = val
        elif field == 7:
            e["slices"] += 1
    return e


def _parse_header(buf):
    h = {"num_shards": 0, "endianness": 0, "version": 0}
    for field, _wt, val in _proto_fields(buf):
        if field == 1:
            h["num_shards"] = val
        elif field == 2:
            h["endianness"] = val
        elif field == 3:
            for f2, _w2, v2 in _proto_fields(val):
                if f2 == 1:
                    h["version"] = v2
    return h


def _tag(field, wt):
    return _put_varint(field << 3 | wt)


def _encode_entry(dtype_id, shape, shard_id, offset, size, masked_crc):
    out = bytearray()
    out += _tag(1, 0) + _put_varint(dtype_id)
    sh = bytearray()
    for d in shape:
        dim = (_tag(1, 0) + _put_varint(int(d))) if d else b""
        sh += _tag(2, 2) + _put_varint(len(dim)) + dim
    out += _tag(2, 2) + _put_varint(len(sh)) + sh
    if shard_id:
        out += _tag(3, 0) + _put_varint(shard_id)
    if offset:
        out += _tag(4, 0) + _put_varint(offset)
    if size:
        out += _tag(5, 0) + _put_varint(size)
    out += _tag(6, 5) + struct.pack("<I", masked_crc)
    return bytes(out)


def _encode_header(num_shards):
    version = _tag(1, 0) + _put_varint(1)  # VersionDef.producer = kTensorBundleVersion
    return _tag(1, 0) + _put_varint(num_shards) + _tag(3, 2) + _put_varint(len(version)) + version


# ----------------------------------------------------------------------------- Snappy (raw format)
def snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            if pos + ln > len(buf):
                raise CheckpointError("snappy: literal runs past the end of the block")
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = (tag >> 5) << 8 | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("snappy: bad copy offset")
        start = len(out) - off
        if off >= ln:
            out += out[start:start + ln]
        else:  # overlapping copy = run-length repetition
            for i in range(ln):
                out.append(out[start + i])
    if len(out) != n:
        raise CheckpointError("snappy: length mismatch (%d vs %d)" % (len(out), n))
    return bytes(out)


# ----------------------------------------------------------------------------- sorted string table
def _read_block(data, offset, size, what):
    end = offset + size
    if offset < 0 or end + BLOCK_TRAILER_LEN > len(data):
        raise CheckpointError("%s block handle points outside the index file" % what)
    raw = data[offset:end]
    ctype = data[end]
    stored = struct.unpack_from("<I", data, end + 1)[0]
    if crc_unmask(stored) != crc32c(data[offset:end + 1]):
        raise CheckpointError("%s block checksum mismatch (corrupt index file)" % what)
    if ctype == NO_COMPRESSION:
        return raw
    if ctype == SNAPPY_COMPRESSION:
        return snappy_uncompress(raw)
    raise CheckpointError("unknown block compression type %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(data):
    """All (key, value) pairs of a sorted string table held in `data` (bytes), in key order."""
    if len(data) < FOOTER_LEN:
        raise CheckpointError("index file shorter than a table footer")
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError("not a TensorFlow checkpoint index (bad table magic)")
    _mo, pos = _get_varint(footer, 0)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    out = []
    for _key, handle in _block_entries(_read_block(data, io, isz, "index")):
        bo, p = _get_varint(handle, 0)
        bs, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, bo, bs, "data")))
    return out


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count % RESTART_INTERVAL == 0:
            if self.count:
                self.restarts.append(len(self.buf))
        else:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))


def write_table(items, block_size=BLOCK_SIZE):
    """Serialise sorted (key, value) byte pairs as an uncompressed sorted string table."""
    out = bytearray()

    def emit(block):
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block)
        out.append(NO_COMPRESSION)
        out.extend(struct.pack("<I", crc_mask(crc32c(block + bytes([NO_COMPRESSION])))))
        return handle

    index, cur, prev = _BlockBuilder(), _BlockBuilder(), None
    for key, value in items:
        if prev is not None and not key > prev:
            raise CheckpointError("table keys must be strictly increasing")
        cur.add(key, value)
        prev = key
        if cur.size() >= block_size:
            index.add(cur.last, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count or not index.count:
        index.add(cur.last, emit(cur.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    return bytes(out)


# ----------------------------------------------------------------------------- the bundle
def _data_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def checkpoint_exists(prefix):
    return os.path.isfile(prefix + ".index")


def list_variables(prefix):
    """{name: (numpy dtype or 'string', shape)} without reading tensor data."""
    with open(prefix + ".index", "rb") as f:
        items = read_table(f.read())
    out = {}
    for key, value in items:
        if key == b"":
            continue
        e = _parse_entry(value)
        out[key.decode("utf-8")] = ("string" if e["dtype"] == DT_STRING else DTYPES.get(e["dtype"]), e["shape"])
    return out


def read_checkpoint(prefix, names=None, verify=True):
    """Read variables of a TF-V2 checkpoint into {name: ndarray}.
    prefix: path without the .index / .data-* suffix (what `tf.train.Saver.restore` takes, src/AE.py:175).
    names:  iterable of variable names to read (default: all numeric tensors); a name that is not in the
            checkpoint raises KeyError, as `Saver.restore` fails with NotFoundError."""
    if not checkpoint_exists(prefix):
        raise FileNotFoundError(prefix + ".index")
    with open(prefix + ".index", "rb") as f:
        items = read_table(f.read())
    if not items or items[0][0] != b"":
        raise CheckpointError("checkpoint index has no bundle header")
    header = _parse_header(items[0][1])
    if header["endianness"] != 0:
        raise CheckpointError("big-endian checkpoints are not supported")
    if header["num_shards"] < 1:
        raise CheckpointError("bundle header declares %d shards" % header["num_shards"])
    entries = {k.decode("utf-8"): _parse_entry(v) for k, v in items[1:]}
    wanted = list(entries) if names is None else list(names)
    missing = [n for n in wanted if n not in entries]
    if missing:
        raise KeyError("%d variable(s) not found in checkpoint %s, e.g. %s" % (len(missing), prefix, missing[:3]))
    shards = {}
    out = {}
    for name in wanted:
        e = entries[name]
        if e["dtype"] == DT_STRING:
            if names is None:
                continue
            raise CheckpointError("%s is a string tensor" % name)
        if e["slices"]:
            raise CheckpointError("%s is a partitioned variable (tensor slices are not supported)" % name)
        dt = DTYPES.get(e["dtype"])
        if dt is None:
            raise CheckpointError("%s has unsupported dtype id %d" % (name, e["dtype"]))
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError("%s: %d bytes stored for shape %s of %s" % (name, e["size"], e["shape"], dt))
        sid = e["shard_id"]
        if sid not in shards:
            if not 0 <= sid < header["num_shards"]:
                raise CheckpointError("%s lives in shard %d of %d" % (name, sid, header["num_shards"]))
            path = _data_path(prefix, sid, header["num_shards"])
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) > 0 \
                else np.zeros(0, np.uint8)  # an empty shard (only zero-sized tensors) cannot be mapped
        blob = shards[sid]
        if e["offset"] < 0 or e["offset"] + e["size"] > blob.size:
            raise CheckpointError("%s runs past the end of its data shard" % name)
        raw = np.array(blob[e["offset"]:e["offset"] + e["size"]])
        if verify and e["crc32c"] is not None and crc_unmask(e["crc32c"]) != crc32c(raw):
            raise CheckpointError("%s: tensor checksum mismatch (corrupt data file)" % name)
        out[name] = raw.view(dt).reshape(e["shape"]).copy()
    return out


def write_checkpoint(prefix, variables):
    """Write {name: ndarray} as a single-shard TF-V2 checkpoint (what `Saver.save` produces, src/AE.py:154-156)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _encode_header(1))]
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in sorted(variables, key=lambda s: s.encode("utf-8")):
            a = np.asarray(variables[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            dtype_id = DTYPE_IDS.get(np.dtype(dt))
            if dtype_id is None:
                raise CheckpointError("%s: dtype %s cannot be stored" % (name, a.dtype))
            raw = np.ascontiguousarray(a.astype(dt, copy=False)).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"),
                          _encode_entry(dtype_id, a.shape, 0, offset, len(raw), crc_mask(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        f.write(write_table(items))
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
