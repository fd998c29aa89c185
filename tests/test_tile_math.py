"""Pins the rank algebra the kernels use (blazeseq_b200/csrc/tile_math.h) on the CPU.

oracle/tile_model.cpp walks arbitrary byte ranges ("runs") with only the scanned prefix as context,
exactly like a CTA (or a GPU shard) does, and must reproduce the oracle's records, SoA destinations
and totals for every run size and window begin.
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REC = np.dtype([(n, "<u4") for n in ("header_start", "seq_start", "sep_start", "qual_start",
                                      "record_end", "seq_dst", "qual_dst", "id_dst", "code")])


@pytest.fixture(scope="module")
def tm(oracle):
    oracle.build()
    L = C.CDLL(os.path.join(ROOT, "oracle", "libbsq_tile_model.so"))
    L.tm_parse.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64,
                           C.c_void_p]
    L.tm_parse.restype = C.c_int64
    L.tm_is_space.argtypes = [C.c_uint32]
    L.tm_check_flags.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32]
    L.tm_check_flags.restype = C.c_int64
    return L


def _rand_stream(rng, nrec, mutate):
    recs = []
    for i in range(nrec):
        L = int(rng.integers(0, 40))
        idl = int(rng.integers(0, 12))
        ident = bytes(rng.choice(list(b"abcXYZ019_ /:"), idl).astype(np.uint8))
        seq = bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8))
        qual = bytes(rng.integers(33, 127, L).astype(np.uint8))
        plus = b"+" + (ident if rng.random() < 0.2 else b"")
        nl = b"\r\n" if mutate == "crlf" else b"\n"
        recs.append(b"@" + ident + nl + seq + nl + plus + nl + qual + nl)
    data = bytearray(b"".join(recs))
    if mutate == "noise" and len(data) > 8:
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(data)))
            data[pos] = int(rng.choice(list(b"\n@+A!\x80\xff ")))
    if mutate == "drop" and len(data) > 8:
        pos = int(rng.integers(0, len(data)))
        del data[pos:pos + int(rng.integers(1, 30))]
    if mutate == "blank":
        data += b"\n" * int(rng.integers(1, 6))
    if mutate == "notail" and data:
        data = data[:-1]
    return bytes(data)


def _check(tm, oracle, data, begin_pad, run_bytes):
    buf = np.frombuffer(b"\xaa" * begin_pad + data, np.uint8)  # bytes before `begin` are foreign
    n = buf.size
    recs = np.zeros(n // 4 + 2, REC)
    tot = np.zeros(8, np.uint32)
    nrec = tm.tm_parse(buf.ctypes.data, begin_pad, n, run_bytes, recs.ctypes.data, recs.size,
                       tot.ctypes.data)
    views, bases, err = oracle.parse_all(data, oracle.config(buffer_growth_enabled=True))
    nl_total = data.count(b"\n")
    assert tot[0] == nl_total and tot[1] == nl_total // 4 == nrec
    k = len(views)
    full = min(k, nrec)  # a tail record without '\n' (Q1) is not a 4-newline record
    v = views[:full]
    r = recs[:full]
    for f in ("header_start", "seq_start", "sep_start", "qual_start", "record_end"):
        assert np.array_equal(r[f].astype(np.int64) - begin_pad, v[f]), f
    assert not r["code"].any()

    def excl(a):
        return np.concatenate([[0], np.cumsum(a)[:-1]])[:full] if full else np.zeros(0, np.int64)

    assert np.array_equal(r["seq_dst"], excl(v["seq_len"]))
    assert np.array_equal(r["qual_dst"], excl(v["qual_len"]))
    raw_id = v["seq_start"] - v["header_start"] - 2
    assert np.array_equal(r["id_dst"], excl(raw_id))
    if err.code in (1, 2, 3):
        # the first structure error is the next 4-newline record, with the same code and position
        assert nrec > k and recs[k]["code"] == err.code
        assert int(recs[k]["header_start"]) - begin_pad == err.file_position
    else:
        assert nrec <= k, "oracle stopped before the 4-newline records ran out"
    if nrec == k:  # totals over complete records
        assert tot[3] == int(v["seq_len"].sum()) and tot[4] == int(v["qual_len"].sum())
        assert tot[5] == int(raw_id.sum())
        consumed = int(v["record_end"][-1]) + 1 if k else 0
        assert int(tot[2]) - begin_pad == consumed


@pytest.mark.parametrize("mutate", ["none", "crlf", "noise", "drop", "blank", "notail"])
def test_runs_reproduce_oracle(tm, oracle, mutate):
    rng = np.random.default_rng(abs(hash(mutate)) % 2**32)
    for trial in range(60):
        data = _rand_stream(rng, int(rng.integers(0, 25)), mutate)
        for run_bytes in (1, 2, 3, 7, 16, 64, 1 << 20, int(rng.integers(1, 200))):
            _check(tm, oracle, data, int(rng.integers(0, 40)), run_bytes)


def test_all_newlines_and_no_newlines(tm, oracle):
    for data in (b"\n" * 257, b"A" * 300, b"", b"\n", b"@\n\n+\n\n" * 50):
        for run_bytes in (1, 5, 64, 4096):
            _check(tm, oracle, data, 3, run_bytes)


def test_synthetic_with_tile_sized_runs(tm, oracle):
    data = oracle.synth(2000, 75, 300, 2, 40, "illumina_1.8").tobytes()
    for run_bytes in (32768, 4096, 319):
        _check(tm, oracle, data, 0, run_bytes)


def test_is_space_and_byte_lane_flags(tm):
    spaces = {9, 10, 11, 12, 13, 28, 29, 30, 32}  # utils.mojo:266-289
    assert {c for c in range(256) if tm.tm_is_space(c)} == spaces
    rng = np.random.default_rng(7)
    words = rng.integers(0, 2**32, 200000, dtype=np.uint64).astype(np.uint32)
    b = words.view(np.uint8)  # bias towards the interesting bytes
    sel = rng.random(b.size) < 0.5
    b[sel] = rng.choice([10, 9, 11, 32, 33, 58, 59, 63, 64, 65, 66, 126, 127, 128, 255, 0],
                        int(sel.sum()))
    for lo, up in ((33, 126), (59, 126), (64, 126), (66, 126)):
        assert tm.tm_check_flags(words.ctypes.data, words.size, lo, up) == 0


@pytest.mark.parametrize("mutate", ["none", "noise", "blank", "notail"])
def test_decoupled_lookback_algebra(tm, mutate):
    """Ordered tree reductions of lb_combine (tile aggregates combined 32 per round back to the nearest
    inclusive state, or to the window-init state) give every tile the prefix of the sequential scan,
    whatever the tile size and whichever predecessors are inclusive: the monoid k_scan_runs relies on."""
    tm.tm_lookback_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    tm.tm_lookback_check.restype = C.c_int64
    rng = np.random.default_rng(abs(hash("lb" + mutate)) % 2**32)
    for trial in range(30):
        data = _rand_stream(rng, int(rng.integers(0, 60)), mutate)
        pad = int(rng.integers(0, 40))
        buf = np.frombuffer(b"\xaa" * pad + data, np.uint8)
        for tile in (1, 3, 16, 61, 4096):      # tile = 1: hundreds of tiles, i.e. many groups of 32
            for inc_every in (0, 1, 2, 5, 32, 33, 40, 1000):
                assert tm.tm_lookback_check(buf.ctypes.data, pad, buf.size, tile, inc_every) == 0, (trial, tile, inc_every)
    for data in (b"\n" * 300, b"A" * 300, b"@\n\n+\n\n" * 50):
        buf = np.frombuffer(data, np.uint8)
        for tile in (1, 2, 7, 64):
            for inc_every in (0, 3, 32):
                assert tm.tm_lookback_check(buf.ctypes.data, 0, buf.size, tile, inc_every) == 0
