"""RapidgzipReader (blazeseq/io/readers.mojo:380-443): the library's parallel gzip decoder (csrc/bsq_pgzip.h, bsq_gzip_*)
must deliver exactly the bytes zlib delivers -- DEFLATE is a lossless standard, so zlib is the oracle -- for every block
type, for streams cut into many speculative chunks, for several members, and must refuse damaged streams.  Host-only:
no GPU is involved (the stream pipeline's use of the same decoder is covered by tests/test_gpu_parity.py)."""
import gzip
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import blazeseq_b200 as B  # noqa: E402
import oracle_py as O  # noqa: E402

CHUNK = 64 << 10   # the smallest speculative chunk: a few MB of FASTQ become dozens of chunks


def read_all(path, parallelism, chunk_bytes=CHUNK, piece=1 << 20):
    r = B.RapidgzipReader(path, parallelism, chunk_bytes=chunk_bytes)
    buf = np.empty(piece, np.uint8)
    out = []
    while True:
        k = r.read_to_buffer(buf, piece, 0)
        if k == 0:
            break
        out.append(buf[:k].tobytes())
    assert r.read_to_buffer(buf, piece, 0) == 0     # stays at the end
    r.close()
    return b"".join(out)


@pytest.fixture(scope="module")
def fastq():
    return O.synth(20000, 100, 200, 2, 40, "sanger").tobytes()     # ~6.6 MB, mixed lengths


def _write(tmp_path, name, blob):
    p = os.path.join(tmp_path, name)
    with open(p, "wb") as f:
        f.write(blob)
    return p


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_fastq_gzip_levels(tmp_path, fastq, level, threads):
    p = _write(tmp_path, "a.fastq.gz", gzip.compress(fastq, level))
    assert read_all(p, threads) == fastq


def test_reference_fixtures(golden_dir):
    """the reference's own .gz / .bgz fixtures (tests/test_data/fastq_parser)"""
    corpus = os.path.join(golden_dir, "corpus")
    n = 0
    for name in sorted(os.listdir(corpus)):
        if name.endswith((".gz", ".bgz")):
            p = os.path.join(corpus, name)
            assert read_all(p, 4) == gzip.open(p).read(), name
            n += 1
    assert n >= 2


def test_block_types_and_members(tmp_path, fastq):
    rng = np.random.default_rng(7)
    noise = rng.integers(0, 256, 1_500_000, dtype=np.uint8).tobytes()
    cases = {}
    cases["stored"] = gzip.compress(fastq[:2_000_000], 0)                      # BTYPE 00 only
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_FIXED)
    cases["fixed"] = co.compress(fastq[:1_000_000]) + co.flush()              # BTYPE 01 only
    cases["noise"] = gzip.compress(noise, 6)                                  # stored blocks chosen by zlib
    runs = b"\0" * 700_000 + b"ab" * 300_000 + b"abc" * 200_000 + fastq[:500_000] + b"x" * 70_000
    cases["runs"] = gzip.compress(runs, 9)                                    # distance 1/2/3 overlapping copies, 258-byte matches
    cases["members"] = b"".join(gzip.compress(fastq[i:i + 700_001], 6) for i in range(0, len(fastq), 700_001))
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, 3_000_000, 50_000):                                     # empty stored blocks between the pieces (pigz-like)
        parts += [co.compress(fastq[i:i + 50_000]), co.flush(zlib.Z_SYNC_FLUSH if (i // 50_000) % 2 else zlib.Z_FULL_FLUSH)]
    parts.append(co.flush())
    cases["flushes"] = b"".join(parts)
    cases["mixed"] = cases["stored"] + cases["fixed"] + gzip.compress(fastq, 6) + gzip.compress(b"") + cases["noise"]
    cases["tiny"] = gzip.compress(b"@r\nACGT\n+\nIIII\n")
    cases["empty"] = gzip.compress(b"")
    cases["named"] = gzip.compress(fastq[:100_000])[:3] + b"\x1c" + gzip.compress(fastq[:100_000])[4:10] + \
        b"\x03\x00abc" + b"name.fq\0" + b"a comment\0" + gzip.compress(fastq[:100_000])[10:]   # FEXTRA + FNAME + FCOMMENT
    for name, blob in cases.items():
        p = _write(tmp_path, name + ".gz", blob)
        want = gzip.open(p).read()
        for threads in (1, 4):
            assert read_all(p, threads) == want, (name, threads)
        assert read_all(p, 4, chunk_bytes=0) == want, name                    # the default 2 MiB chunks (one or two chunks here)


def test_small_reads_and_positions(tmp_path, fastq):
    p = _write(tmp_path, "b.fastq.gz", gzip.compress(fastq, 6))
    r = B.RapidgzipReader(p, 4, chunk_bytes=CHUNK)
    buf = np.zeros(10_000, np.uint8)
    got = bytearray()
    sizes = [1, 7, 4096, 9999, 33]
    i = 0
    while True:
        amt = sizes[i % len(sizes)]
        pos = (i * 13) % (buf.size - amt + 1)
        k = r.read_to_buffer(buf, amt, pos)          # Reader.read_to_buffer(buf, amt, pos), readers.mojo:421-443
        if k == 0:
            break
        got += buf[pos:pos + k].tobytes()
        i += 1
    assert bytes(got) == fastq
    with pytest.raises(B.BlazeSeqError):
        r.read_to_buffer(buf, buf.size + 1, 0)
    with pytest.raises(B.BlazeSeqError):
        r.read_to_buffer(buf, 10, buf.size + 1)
    r.close()


def test_damaged_streams_are_refused(tmp_path, fastq):
    blob = bytearray(gzip.compress(fastq, 6))
    # a flipped bit in the middle of the deflate data: either the codes stop making sense or the CRC-32 differs
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x10
    p = _write(tmp_path, "flip.gz", bytes(bad))
    with pytest.raises(B.BlazeSeqError):
        read_all(p, 4)
    # wrong CRC-32 in the trailer
    bad = bytearray(blob)
    bad[-8] ^= 0xFF
    p = _write(tmp_path, "crc.gz", bytes(bad))
    with pytest.raises(B.BlazeSeqError):
        read_all(p, 4)
    # wrong ISIZE
    bad = bytearray(blob)
    bad[-1] ^= 0x01
    p = _write(tmp_path, "isize.gz", bytes(bad))
    with pytest.raises(B.BlazeSeqError):
        read_all(p, 4)
    # truncated
    p = _write(tmp_path, "cut.gz", bytes(blob[: len(blob) * 2 // 3]))
    with pytest.raises(B.BlazeSeqError):
        read_all(p, 4)
    # not gzip at all
    p = _write(tmp_path, "plain.gz", fastq[:1000])
    with pytest.raises(OSError):
        read_all(p, 4)
    with pytest.raises(OSError):
        B.RapidgzipReader(os.path.join(tmp_path, "missing.gz"), 4)


def test_trailing_garbage_is_ignored_like_gzip(tmp_path, fastq):
    p = _write(tmp_path, "trail.gz", gzip.compress(fastq[:300_000]) + b"\0" * 512)
    assert read_all(p, 4) == fastq[:300_000]


def test_random_streams_differential(tmp_path):
    """Seeded random payloads x compression level / strategy / flush pattern x chunk size x threads: always zlib's bytes."""
    rng = np.random.default_rng(20260)
    alphabets = [b"ACGT", b"ACGTN\n@+!#$%&'()*+,-./0123456789:;<=>?", bytes(range(256)), b"a"]
    for case in range(24):
        alpha = alphabets[case % len(alphabets)]
        n = int(rng.integers(1, 1_200_000))
        idx = rng.integers(0, len(alpha), n)
        payload = bytes(np.frombuffer(alpha, np.uint8)[idx])
        if case % 3 == 0:                         # long repeats at assorted distances
            piece = payload[: int(rng.integers(1, 5000))]
            payload = (piece * (n // len(piece) + 1))[:n]
        level = int(rng.integers(0, 10))
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][case % 5]
        co = zlib.compressobj(level, zlib.DEFLATED, 31, int(rng.integers(1, 10)), strategy)
        parts, pos = [], 0
        while pos < n:
            step = int(rng.integers(1, 400_000))
            parts.append(co.compress(payload[pos:pos + step]))
            if rng.random() < 0.3:
                parts.append(co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH][int(rng.integers(0, 2))]))
            pos += step
        parts.append(co.flush())
        blob = b"".join(parts)
        if case % 4 == 1:
            blob = blob + gzip.compress(payload[: n // 3], int(rng.integers(1, 10)))      # a second member
        p = _write(tmp_path, "r%d.gz" % case, blob)
        want = gzip.open(p).read()
        chunk = [CHUNK, 100_000, 333_333, 0][case % 4]
        threads = [1, 2, 5, 8][case % 4]
        assert read_all(p, threads, chunk_bytes=chunk, piece=int(rng.integers(1, 1 << 20))) == want, (case, level, strategy, chunk, threads)


def test_corrupted_streams_fuzz(tmp_path, fastq):
    """Seeded corruption of a multi-chunk stream (bit flips, overwritten runs, truncations, inserted garbage): the decoder
    either refuses the stream or -- when the damage is not in bytes that matter -- delivers exactly what zlib delivers;
    it never hangs and never hands out bytes of a stream whose CRC-32 does not hold."""
    rng = np.random.default_rng(4242)
    blob = gzip.compress(fastq[:3_000_000], 6) + gzip.compress(fastq[3_000_000:4_000_000], 1)
    for case in range(40):
        bad = bytearray(blob)
        kind = case % 4
        at = int(rng.integers(12, len(bad) - 12))
        if kind == 0:
            bad[at] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            n = int(rng.integers(1, 2000))
            bad[at:at + n] = rng.integers(0, 256, min(n, len(bad) - at), dtype=np.uint8).tobytes()
        elif kind == 2:
            del bad[at:]
        else:
            bad[at:at] = rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8).tobytes()
        p = _write(tmp_path, "f%d.gz" % case, bytes(bad))
        try:
            want = gzip.open(p).read()
        except Exception:
            want = None
        try:
            got = read_all(p, 4)
        except B.BlazeSeqError:
            got = None
        if got is not None:
            assert want is not None and got == want, (case, kind, at)
