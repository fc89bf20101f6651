"""A byte stream compressed block by block into packed block containers (the
reference has no serialised form besides Show/Read, src/Data/RLE/Internal.hs:95-96).

Layout (little endian):  b"TCZ1" | u32 block_bytes | u64 total_bytes | u64 n_blocks |
                         n_blocks x ( u64 container_bytes | container )
Every container is what tc_blocks_encode_packed writes for one block (include/tc_b200.h,
tc_packed_header); blocks are independent, so a multi-GPU writer shards them round-robin
(multi.compress_blocks_sharded(..., packed=True)) and concatenates in block order.
"""
from __future__ import annotations

import struct

import numpy as np

from . import block

MAGIC = b"TCZ1"
_HEAD = struct.Struct("<4sIQQ")


def compress_stream(data, block_bytes: int = 16 << 20, with_mtf: bool = True, ctx=None) -> bytes:
    """bytes -> TCZ1 stream.  All blocks go through ONE tc_blocks_encode_packed call (copies
    overlapped with the kernels, several blocks in flight)."""
    if block_bytes <= 0 or block_bytes >= (1 << 32) - 4:
        raise ValueError("block_bytes must be in 1 .. 2^32 - 5")
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    blocks = [buf[o:o + block_bytes] for o in range(0, buf.size, block_bytes)]
    blobs = block.compress_blocks_packed(blocks, with_mtf, ctx) if blocks else []
    out = [_HEAD.pack(MAGIC, block_bytes, buf.size, len(blobs))]
    for b in blobs:
        out.append(struct.pack("<Q", b.size))
        out.append(b.tobytes())
    return b"".join(out)


def split_stream(blob):
    """TCZ1 stream -> (block_bytes, total_bytes, [container bytes]); raises ValueError if malformed."""
    blob = bytes(blob)
    if len(blob) < _HEAD.size:
        raise ValueError("TCZ1: truncated header")
    magic, block_bytes, total, nb = _HEAD.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("TCZ1: bad magic")
    off = _HEAD.size
    parts = []
    for _ in range(nb):
        if off + 8 > len(blob):
            raise ValueError("TCZ1: truncated block table")
        (sz,) = struct.unpack_from("<Q", blob, off)
        off += 8
        if off + sz > len(blob):
            raise ValueError("TCZ1: truncated container")
        parts.append(blob[off:off + sz])
        off += sz
    return block_bytes, total, parts


def decompress_stream(blob, ctx=None) -> bytes:
    """TCZ1 stream -> bytes (all containers through one tc_blocks_decode_packed call: several blocks in flight)."""
    _, total, parts = split_stream(blob)
    out = b"".join(block.decompress_blocks_packed([np.frombuffer(p, dtype=np.uint8) for p in parts], ctx))
    if len(out) != total:
        raise ValueError(f"TCZ1: decoded {len(out)} bytes, header says {total}")
    return out
