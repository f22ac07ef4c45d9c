"""Container of one image's PC1 entropy-coded symbols (host side; the payload format is csrc/pc_codec.cu).

    offset  size  field
    0       4     magic "DSPC"
    4       1     version (1)
    5       1     L   number of quantiser centres
    6       2     C   bottleneck channels (depth of the symbol volume), little endian
    8       2     H   symbol rows
    10      2     W   symbol columns
    12      1     S   number of range-coder streams (depth slice d is in stream d % S)
    13      3     reserved, 0
    16      4*S   stream sizes in bytes, little endian
    ...           the S streams, concatenated

The reference never serialises symbols (its helpers stop at per-symbol frequencies,
src/probclass_imgcomp.py:361-482), so this layout has no counterpart there.
"""
from __future__ import annotations

import struct

MAGIC = b"DSPC"
VERSION = 1
_HEAD = struct.Struct("<4sBBHHHB3x")


def pack(streams, c, h, w, L):
    if not 1 <= len(streams) <= 255:
        raise ValueError("1..255 streams")
    head = _HEAD.pack(MAGIC, VERSION, L, c, h, w, len(streams))
    sizes = struct.pack("<%dI" % len(streams), *[len(s) for s in streams])
    return head + sizes + b"".join(streams)


def unpack(blob):
    """-> (C, H, W, L, [stream bytes])."""
    if len(blob) < _HEAD.size:
        raise ValueError("bitstream shorter than its header")
    magic, version, L, c, h, w, ns = _HEAD.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("not a DSIN PC1 bitstream (bad magic)")
    if version != VERSION:
        raise ValueError("unsupported bitstream version %d" % version)
    if ns < 1 or len(blob) < _HEAD.size + 4 * ns:
        raise ValueError("truncated stream table")
    sizes = struct.unpack_from("<%dI" % ns, blob, _HEAD.size)
    pos = _HEAD.size + 4 * ns
    if pos + sum(sizes) != len(blob):
        raise ValueError("bitstream length does not match its stream table")
    streams = []
    for s in sizes:
        streams.append(bytes(blob[pos:pos + s]))
        pos += s
    return c, h, w, L, streams


def payload_bits(blob):
    """Bits of entropy-coded payload (without the container header)."""
    _c, _h, _w, _L, streams = unpack(blob)
    return 8 * sum(len(s) for s in streams)
