"""BGZF (blocked gzip, SAM specification 4.1) writer used by the tests and the gzip bench.

A BGZF file is a series of gzip members of at most 64 KiB, each carrying its compressed size in a 'BC'
extra field, closed by an empty member.  The native stream pipeline (bsq_stream_*) inflates such
members block-parallel; any gzip reader concatenates them."""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
BLOCK = 0xFF00


def _member(chunk: bytes, level: int) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    d = c.compress(chunk) + c.flush()
    assert len(d) + 26 <= 0x10000
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(d) + 25) + d +
            struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def compress(data, level: int = 6, threads: int = 1) -> bytes:
    data = memoryview(data).cast("B")
    chunks = [bytes(data[i:i + BLOCK]) for i in range(0, len(data), BLOCK)]
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda ch: _member(ch, level), chunks, chunksize=64))
    else:
        parts = [_member(ch, level) for ch in chunks]
    return b"".join(parts) + _EOF
