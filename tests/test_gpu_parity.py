"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same bytes.

Bit-exact bar: record count, the five RecordOffsets of every record, the stripped id span, the
FastqBatch SoA arrays (bytes and Int64 ends per batch), totals, and the stop reason with the
reference's context and message text.  Run with `pytest -m gpu` on the B200 box.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from ref_cases import BATCH_CASES, CASES, EXAMPLE_IDS  # noqa: E402


@pytest.fixture(scope="module")
def B():
    import blazeseq_b200
    return blazeseq_b200


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


def _cfg_kw(kw):
    return dict(check_ascii=kw.get("check_ascii", False), check_quality=kw.get("check_quality", False),
                schema=kw.get("schema", "generic"), growth=kw.get("buffer_growth_enabled", False),
                buffer_capacity=kw.get("buffer_capacity"), buffer_max_capacity=kw.get("buffer_max_capacity"))


# ------------------------------------------------------------------ reference literal streams


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("api", ["next_view", "next_record"])
def test_literal_streams_through_fastq_parser(B, oracle, case, api):
    """The reference's own unit tests, run against the drop-in API."""
    name, cite, data, kw, records, err_sub = case
    cfg = B.ParserConfig(check_ascii=kw.get("check_ascii", False), check_quality=kw.get("check_quality", False),
                         buffer_capacity=kw.get("buffer_capacity", B.DEFAULT_CAPACITY),
                         buffer_growth_enabled=kw.get("buffer_growth_enabled", False),
                         buffer_max_capacity=kw.get("buffer_max_capacity", B.MAX_CAPACITY))
    p = B.FastqParser(B.MemoryReader(data), kw.get("schema"), config=cfg)
    for exp in records:
        r = getattr(p, api)()
        got = (r.id(), r.sequence(), r.quality()) if api == "next_view" else (r._id, r._sequence, r._quality)
        assert got == exp, cite
    with pytest.raises(B.BlazeSeqError) as ei:
        getattr(p, api)()
    assert err_sub in str(ei.value), (cite, str(ei.value))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_literal_streams_bit_exact(U, oracle, case):
    name, cite, data, kw, records, err_sub = case
    U.check_stream(oracle, data, batch_size=2, **_cfg_kw(kw))


@pytest.mark.parametrize("cite,data,bs,sizes", BATCH_CASES, ids=[c[0] for c in BATCH_CASES])
def test_batches_api(B, cite, data, bs, sizes):
    p = B.FastqParser(B.MemoryReader(data), batch_size=bs, schema="generic")
    got = [len(b) for b in p.batches()]
    assert got == sizes, cite
    assert not p.has_more()
    p = B.FastqParser(B.MemoryReader(data), batch_size=bs, schema="generic")
    got = []
    for _ in range(len(sizes)):
        got.append(len(p.next_batch(bs)))
    assert got == sizes


def test_batch_content_and_layout(B):
    """tests/fastq/test_parser.mojo:163-177, tests/fastq/test_record_batch.mojo:26-38."""
    p = B.FastqParser(B.MemoryReader(b"@seq1\nACGT\n+\n!!!!\n"), batch_size=4, schema="generic")
    batch = p.next_batch(4)
    rec = batch.get_record(0)
    assert (rec.id, rec.sequence, rec.quality) == ("seq1", "ACGT", "!!!!")
    p = B.FastqParser(B.MemoryReader(b"@a\nAC\n+\n!!\n@b\nGT\n+\n!!\n"), batch_size=4, schema="generic")
    batch = p.next_batch(4)
    assert batch._ends.tolist() == [2, 4] and batch.seq_len() == 4 and batch.num_records() == 2
    assert batch.to_device() is not None and batch.to_device().num_records == 2


def test_python_binding_surface(B, golden_dir):
    """tests/test_python_bindings.py:31-118 against blazeseq_b200.parser()."""
    path = os.path.join(golden_dir, "corpus", "example.fastq")
    p = B.parser(path, "generic")
    assert p.has_more()
    count = 0
    while True:
        try:
            rec = p.next_record()
            count += 1
            if count == 1:
                assert rec.id == EXAMPLE_IDS[0].decode()
                assert "CCCTTCTTGTCTTCAGCGTTTCTCC" in rec.sequence
                assert len(rec) == len(rec.sequence) and len(rec.phred_scores) >= len(rec)
        except Exception as e:
            assert "EOF" in str(e)
            break
    assert count == 3
    p = B.parser(path, "generic")
    b1 = p.next_batch(2)
    assert b1.num_records() == 2 and [r.id for r in b1] == [x.decode() for x in EXAMPLE_IDS[:2]]
    b2 = p.next_batch(10)
    assert b2.num_records() == 1 and b2.get_record(0).id == EXAMPLE_IDS[2].decode()
    assert [r.id for r in B.parser(path, "generic")] == [x.decode() for x in EXAMPLE_IDS]
    gz = B.parser(os.path.join(golden_dir, "corpus", "example.fastq.gz"), "generic")
    assert [r.id for r in gz.records()] == [x.decode() for x in EXAMPLE_IDS]
    bgz = B.parser(os.path.join(golden_dir, "corpus", "example.fastq.bgz"), "generic")
    assert len(list(bgz.records())) == 3


def test_iterator_swallows_error(B, capsys):
    """tests/test_error_context.mojo:81-94: the iterator prints the error and yields nothing."""
    from ref_cases import INVALID_ID
    p = B.FastqParser(B.MemoryReader(INVALID_ID), config=B.ParserConfig(check_ascii=True, check_quality=True))
    assert list(p.records()) == []
    assert "Record number: 1" in capsys.readouterr().out


# ------------------------------------------------------------------ corpus


def _corpus(golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "corpus_expect.json")))
    for f in sorted(exp["files"]):
        yield f, exp["files"][f], open(os.path.join(golden_dir, "corpus", f), "rb").read()


def test_corpus_validation_on(U, B, oracle, golden_dir):
    """All 70 files under the invalid-test config (check_ascii, check_quality, generic schema)."""
    gpu = B.GpuParser(True, True, B.parse_schema("generic"), 3)
    for f, row, data in _corpus(golden_dir):
        U.check_stream(oracle, data, check_ascii=True, check_quality=True, batch_size=3, gpu=gpu)
    gpu.close()


def test_corpus_validation_off_file_schema(U, B, oracle, golden_dir):
    """Valid files with their own schema, validation off (the reference's valid-file tests), and
    again with the quality check on for that schema."""
    for f, row, data in _corpus(golden_dir):
        schema = row["valid_schema"] or "generic"
        U.check_stream(oracle, data, schema=schema, batch_size=4096)
        U.check_stream(oracle, data, schema=schema, check_quality=True, batch_size=7, growth=True)


def test_corpus_readme_expectations(B, golden_dir):
    """The stop message must contain what the reference's own tests accept
    (tests/fastq/test_fastq_parser_correctness.mojo:21-56)."""
    exp = json.load(open(os.path.join(golden_dir, "corpus_expect.json")))
    for f, row, data in _corpus(golden_dir):
        if not row["invalid_msg"]:
            continue
        p = B.FastqParser(B.MemoryReader(data), config=B.ParserConfig(check_ascii=True, check_quality=True))
        with pytest.raises(B.BlazeSeqError) as ei:
            while True:
                p.next_record()
        msg = str(ei.value)
        assert row["invalid_msg"] in msg or any(a in msg for a in exp["accept_set"]), (f, msg)


# ------------------------------------------------------------------ adversarial streams


def _rand_stream(rng, nrec, mutate, maxlen=200):
    recs = []
    for i in range(nrec):
        L = int(rng.integers(0, maxlen))
        idl = int(rng.integers(0, 24))
        ident = bytes(rng.choice(list(b"abcXYZ019_ /:\t"), idl).astype(np.uint8))
        seq = bytes(rng.choice(list(b"ACGTN"), L).astype(np.uint8))
        qual = bytes(rng.integers(33, 127, L).astype(np.uint8))
        plus = b"+" + (ident if rng.random() < 0.2 else b"")
        nl = b"\r\n" if mutate == "crlf" else b"\n"
        recs.append(b"@" + ident + nl + seq + nl + plus + nl + qual + nl)
    data = bytearray(b"".join(recs))
    if mutate == "noise" and len(data) > 8:
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(len(data) // 2, len(data)))
            data[pos] = int(rng.choice(list(b"\n@+A!\x80\xff ")))
    if mutate == "drop" and len(data) > 8:
        pos = int(rng.integers(len(data) // 2, len(data)))
        del data[pos:pos + int(rng.integers(1, 30))]
    if mutate == "blank":
        data += b"\n" * int(rng.integers(1, 6))
    if mutate == "notail" and data:
        data = data[:-1]
    return bytes(data)


@pytest.mark.parametrize("mutate", ["none", "crlf", "noise", "drop", "blank", "notail"])
def test_random_streams_multi_tile(U, B, oracle, mutate):
    """Streams of a few hundred KB: records straddle every kind of tile and run boundary."""
    rng = np.random.default_rng(abs(hash(mutate)) % 2**32)
    for val in (False, True):
        gpu = B.GpuParser(val, val, B.parse_schema("generic"), 64, buffer_growth_enabled=True)
        for trial in range(6):
            data = _rand_stream(rng, int(rng.integers(1, 3000)), mutate)
            # views only / batches only / both: the passes differ (k_summarize<false> vs <true>)
            U.check_stream(oracle, data, check_ascii=val, check_quality=val, batch_size=64, growth=True, gpu=gpu,
                           want=(3, 1, 2)[trial % 3])
        gpu.close()


def test_tiny_and_degenerate_streams(U, oracle):
    for data in (b"", b"\n", b"\n\n\n\n", b"@\n\n+\n\n", b"@a\nA\n+\n!\n", b"A" * 100, b"@a\nAC\n+\n!!",
                 b"@a\nAC\n+\n!!\n@b\nAC\n+\n!", b"@a\nAC\n+\n!!\n\n\n\n", b"@a\nAC\n+\n!!\n  \n",
                 b"@ a \nAC\n+\n!!\n", b"@\t\nAC\n+\n!!\n", b"@   \nAC\n+\n!!\n", b"@a\n\n+\n\n@b\n\n+\n\n"):
        for growth in (False, True):
            U.check_stream(oracle, data, batch_size=1, growth=growth)
            U.check_stream(oracle, data, check_ascii=True, check_quality=True, batch_size=2, growth=growth)
            U.check_stream(oracle, data, batch_size=3, growth=growth, want=1)
            U.check_stream(oracle, data, check_quality=True, batch_size=3, growth=growth, want=2)


def test_all_newlines_overflows_the_tile_list(U, oracle):
    """More newlines in one 16 KiB tile than the newline list holds (512): the resolve kernel walks the tile in several passes."""
    U.check_stream(oracle, b"\n" * 100000, batch_size=16)
    U.check_stream(oracle, b"@\n\n+\n\n" * 30000, batch_size=4096, check_ascii=True, check_quality=True)
    U.check_stream(oracle, b"@ab\nA\n+\nI\n" * 40000, batch_size=1000)


def test_long_reads_span_tiles(U, oracle):
    """Records far longer than a 16 KiB tile (and than a CTA's run)."""
    rng = np.random.default_rng(11)
    recs = []
    for L in (100000, 5, 70000, 32768, 32767, 32769, 1, 250000, 0, 65536):
        seq = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))
        qual = bytes(rng.integers(33, 127, L).astype(np.uint8))
        recs.append(b"@read/%d some description\n" % L + seq + b"\n+\n" + qual + b"\n")
    data = b"".join(recs)
    U.check_stream(oracle, data, batch_size=3)
    U.check_stream(oracle, data, batch_size=4, check_ascii=True, check_quality=True)
    U.check_stream(oracle, data[:-1], batch_size=4, check_ascii=True, check_quality=True)  # Q1 tail
    bad = bytearray(data)
    bad[len(recs[0]) + len(recs[1]) + 40000] = 0x80  # non-ASCII deep inside a long sequence line
    U.check_stream(oracle, bytes(bad), batch_size=4, check_ascii=True)


@pytest.mark.parametrize("growth", [False, True])
def test_record_longer_than_the_buffer(U, B, oracle, growth):
    """parser.mojo:484-503: a record that does not fit buffer_capacity (growth off) / buffer_max_capacity (growth
    on) ends the parse with BUFFER_EXCEEDED / BUFFER_AT_MAX and the reference's text; the records before it are
    delivered.  Lengths right at the limit, the long record first / in the middle / last, with and without a
    trailing newline, next to structure errors in the same record."""
    rng = np.random.default_rng(5 + growth)

    def rec(i, L, bad=False):
        seq = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))
        return b"@r%d\n" % i + seq + b"\n+\n" + (b"I" * (L - 1 if bad else L)) + b"\n"
    cap = 300
    kw = dict(buffer_capacity=cap if not growth else 64, buffer_max_capacity=cap if growth else 1 << 20, growth=growth)
    for long_at in (0, 7, 19):
        for total in (cap - 1, cap, cap + 1, 5 * cap):
            recs = [rec(i, int(rng.integers(20, 100))) for i in range(20)]
            head = len(b"@r%d\n" % long_at) + 4                    # '@id\n' + '\n+\n' + the two line ends
            L = (total - head) // 2
            recs[long_at] = rec(long_at, L)
            if len(recs[long_at]) != total:                        # odd totals: pad the id
                recs[long_at] = recs[long_at].replace(b"@r", b"@rr", 1)
            data = b"".join(recs)
            res = U.check_stream(oracle, data, batch_size=8, **kw)
            if len(recs[long_at]) > cap:
                assert res.n_records == long_at and res.stop.code == (9 if growth else 8), (long_at, total, res.stop.text)
                assert ("maximum buffer capacity (%d bytes)" % cap if growth else "buffer capacity (%d bytes)" % cap) in res.stop.text
            else:
                assert res.n_records == 20
            U.check_stream(oracle, data[:-1], batch_size=8, **kw)                    # no trailing newline
            U.check_stream(oracle, data, batch_size=8, want=2, **kw)
    # too long AND malformed: the buffer error comes first (the record is never scanned to its end)
    recs = [rec(0, 30), rec(1, 400, bad=True), rec(2, 30)]
    res = U.check_stream(oracle, b"".join(recs), batch_size=8, **kw)
    assert res.n_records == 1 and res.stop.code == (9 if growth else 8)
    # an unterminated fragment longer than the buffer
    U.check_stream(oracle, rec(0, 30) + b"@frag\n" + b"A" * 500, batch_size=8, **kw)
    U.check_stream(oracle, rec(0, 30) + b"@frag\n" + b"A" * 200 + b"\n+\n" + b"I" * 200, batch_size=8, **kw)


@pytest.mark.parametrize("width", [16, 32, 64])
def test_quality_check_as_written_in_the_reference(U, B, oracle, width):
    """bsq_config.compat_q5_width = W reproduces Validator._validate_quality_range as written (record.mojo:90-102):
    the first floor(n / W) * W quality bytes of a record are also rejected when they EQUAL the schema's upper bound
    ('~'), the remaining ones only above it.  Oracle: ora_config.compat_simd_width."""
    rng = np.random.default_rng(width)
    base = oracle.synth(4000, 90, 210, 2, 40, "sanger")           # several tiles; no '~' anywhere
    views, _, _ = oracle.parse_all(base, oracle.config(True, True, "sanger"))
    kw = dict(check_quality=True, schema="sanger", batch_size=500, compat_q5_width=width)
    U.check_stream(oracle, base, **kw)
    for trial in range(12):
        k = int(rng.integers(0, len(views)))
        v = views[k]
        n = int(v["qual_len"])
        body = n - n % width
        where = [0, body - 1, body, n - 1, int(rng.integers(0, n))][trial % 5]
        if where < 0 or where >= n:
            continue
        data = base.copy()
        data[int(v["qual_start"]) + where] = ord("~")
        res = U.check_stream(oracle, data, **kw)
        assert (res.n_records == k and res.stop.code == 5) if where < body else res.n_records == len(views), (k, where, body)
        U.check_stream(oracle, data, check_quality=True, schema="sanger", batch_size=500)   # documented intent: '~' is valid
    # a long read whose '~' sits in an earlier tile than the end of its quality line
    L = 40007
    seq = bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8))
    for where in (5, L - L % width - 1, min(L - L % width, L - 1), L - 1, L - 7):
        q = bytearray(b"I" * L)
        q[where] = ord("~")
        data = b"@a\nAC\n+\nII\n@long\n" + seq + b"\n+\n" + bytes(q) + b"\n@b\nAC\n+\nII\n"
        U.check_stream(oracle, data, check_quality=True, schema="sanger", batch_size=2, compat_q5_width=width)


@pytest.mark.parametrize("digits", [8, 9])
def test_record_stride_sweep(U, B, oracle, digits):
    """Constant-stride streams at every record stride 293 ... 354 bytes (2 L + id digits + 11; 320 = 16 banks
    apart, 352 = 8): the SoA copy is organised by destination vector, so no stride is special -- and every
    alignment of (source - destination) mod 16, every line-end position inside a vector, is exercised."""
    gpu = B.GpuParser(False, False, B.parse_schema("generic"), 1000)
    total = 10 ** (digits - 1) + 1                 # ids zero padded to `digits`
    for L in range(137, 168):
        data = oracle.synth(total, L, L, 2, 40, "sanger", first=12345, count=700)
        assert data.size == 700 * (2 * L + digits + 11)
        U.check_stream(oracle, data, batch_size=1000, gpu=gpu, want=2 if L % 2 else 3)
    gpu.close()


def test_validation_screen_of_clean_and_dirty_tiles(U, B, oracle):
    """With validation on, k_summarize screens every interior tile for HI / BAD bytes and k_resolve examines
    only the flagged ones.  One corrupted byte at a time -- in an id, a sequence, a '+' line (HI there is not
    an error), a quality line, next to 16 KiB tile edges -- must give exactly the oracle's verdict, and the
    clean stream none."""
    data = oracle.synth(9000, 100, 200, 2, 40, "sanger")          # ~3 MB, ids without blanks: every tile is clean
    views, bases, err = oracle.parse_all(data, oracle.config(True, True, "sanger"))
    assert len(views) == 9000
    gpu = B.GpuParser(True, True, B.parse_schema("sanger"), 1000)
    U.check_stream(oracle, data, check_ascii=True, check_quality=True, schema="sanger", batch_size=1000, gpu=gpu)
    rng = np.random.default_rng(17)
    tile = 16384
    spots = []
    for k in (3000, 6100):
        v = views[k]
        spots += [(int(v["id_start"]) + 2, 0x80), (int(v["seq_start"]) + 5, 0xC3), (int(v["sep_start"]), 0x80 | ord("+")),
                  (int(v["qual_start"]) + 7, 0xFF), (int(v["qual_start"]) + 1, ord(" ")), (int(v["seq_start"]) + 1, ord(" ")),
                  (int(v["id_start"]) + 1, 0x7F), (int(v["qual_start"]) + 3, 0x7F)]
    for edge in (40 * tile, 41 * tile - 1, 41 * tile, 97 * tile + 1):
        spots += [(edge, 0x80), (edge, 0x1F)]
    for pos, byte in spots:
        if data[pos] == 10:
            continue
        bad = data.copy()
        bad[pos] = byte
        U.check_stream(oracle, bad, check_ascii=True, check_quality=True, schema="sanger", batch_size=1000, gpu=gpu,
                       want=(3, 1, 2)[int(rng.integers(0, 3))])
    for val in ((True, False), (False, True)):                     # each validator on its own
        g2 = B.GpuParser(val[0], val[1], B.parse_schema("sanger"), 1000)
        for pos, byte in spots[3:6]:
            bad = data.copy()
            bad[pos] = byte
            U.check_stream(oracle, bad, check_ascii=val[0], check_quality=val[1], schema="sanger", batch_size=1000, gpu=g2)
        g2.close()
    gpu.close()


def test_device_consumer_quality_sums(B, oracle):
    """bsq_quality_sums reads the device-resident SoA of a batches() pass (quality arena + per-batch ends):
    per-record Phred sums equal the host computation on the input bytes, across batch and window edges."""
    from blazeseq_b200 import _capi as capi
    rng = np.random.default_rng(21)
    data = np.frombuffer(_rand_stream(rng, 5000, "none", maxlen=260), np.uint8)
    views, bases, err = oracle.parse_all(data.tobytes())
    exp = np.array([int(data[int(v["qual_start"]):int(v["qual_start"]) + int(v["qual_len"])].astype(np.int64).sum())
                    - 33 * int(v["qual_len"]) for v in views], np.int64)
    for m in (1, 7, 512, 4096):
        gpu = B.GpuParser(False, False, B.parse_schema("generic"), m, buffer_growth_enabled=True)
        res = gpu.parse_host(data, want=capi.WANT_BATCHES)
        assert res.n_records == len(views)
        got = gpu.quality_sums()
        assert np.array_equal(got.astype(np.int64), exp)
        assert np.array_equal(gpu.quality_sums(1234, 100).astype(np.int64), exp[1234:1334])
        gpu.close()


def test_id_strip_paths_agree(U, oracle):
    """CRLF and padded ids take the strip pipeline; forcing it on clean input changes nothing."""
    rng = np.random.default_rng(5)
    clean = _rand_stream(rng, 2000, "none")
    crlf = _rand_stream(rng, 2000, "crlf")
    for data in (clean, crlf):
        a = U.check_stream(oracle, data, batch_size=100)
        b = U.check_stream(oracle, data, batch_size=100, force_id_slow=True)
        assert b.id_slow_path == 1
    assert U.check_stream(oracle, oracle.synth(3000, 50, 150, 2, 40, "sanger"), batch_size=512).id_slow_path == 0
    assert U.check_stream(oracle, crlf, batch_size=100).id_slow_path == 1


@pytest.mark.parametrize("schema", ["generic", "sanger", "solexa", "illumina_1.3", "illumina_1.5", "illumina_1.8"])
def test_quality_schemas(U, oracle, schema):
    """Every schema's bounds, with qualities that sit exactly on and just outside them."""
    lo, up, off, _ = oracle.schema(schema)
    rng = np.random.default_rng(lo * 7 + up)
    good = []
    for i in range(300):
        L = int(rng.integers(1, 120))
        q = rng.integers(lo, up + 1, L).astype(np.uint8)
        q[0] = lo
        q[-1] = up
        good.append(b"@r%d\n" % i + b"A" * L + b"\n+\n" + bytes(q) + b"\n")
    data = b"".join(good)
    U.check_stream(oracle, data, schema=schema, check_quality=True, check_ascii=True, batch_size=50)
    for badbyte in (lo - 1, up + 1, 0x80, 0xFF, 9, 13):
        mut = bytearray(data)
        # last quality byte of record 150
        pos = sum(len(g) for g in good[:151]) - 2
        mut[pos] = badbyte
        U.check_stream(oracle, bytes(mut), schema=schema, check_quality=True, batch_size=50)
        U.check_stream(oracle, bytes(mut), schema=schema, check_quality=True, check_ascii=True, batch_size=50)


# ------------------------------------------------------------------ BASELINE configs


def test_config1_1k_records_validation_on(U, oracle):
    """BASELINE.json configs[0]: 1k-record 150 bp synthetic, validation ON -- bit-exact."""
    data = oracle.synth(1000, 150, 150, 2, 40, "sanger")
    assert data.size == 314000
    res = U.check_stream(oracle, data, schema="sanger", check_ascii=True, check_quality=True, batch_size=4096)
    assert (res.n_records, res.n_bases) == (1000, 150000)
    res = U.check_stream(oracle, data, schema="sanger", check_ascii=True, check_quality=True, batch_size=4096,
                         via="device")
    assert res.stop.code == 6


def test_mixed_lengths_and_device_input(U, oracle):
    data = oracle.synth(20000, 75, 300, 2, 40, "illumina_1.8")
    for via in ("host", "device"):
        res = U.check_stream(oracle, data, schema="illumina_1.8", batch_size=4096, via=via)
        assert res.n_records == 20000


def test_pageable_and_pinned_host_sources_agree(B, oracle):
    import torch
    data = oracle.synth(50000, 150, 150, 2, 40, "illumina_1.8")
    gpu = B.GpuParser(False, False, B.parse_schema("illumina_1.8"), 4096, h2d_chunk_bytes=1 << 20)
    r1 = gpu.parse_host(data, want=3)
    a = gpu.batch_to_host(5)
    pinned = torch.from_numpy(data.copy()).pin_memory()
    r2 = gpu.parse_host(pinned.numpy(), want=3)
    b = gpu.batch_to_host(5)
    assert (r1.n_records, r1.n_bases) == (r2.n_records, r2.n_bases) == (50000, 7500000)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    gpu.close()


# ------------------------------------------------------------------ synthetic generator on the device


def test_device_generator_matches_reference_generator(B, oracle, golden_dir):
    import torch
    kat = json.load(open(os.path.join(golden_dir, "synthetic_kat.json")))
    gpu = B.GpuParser()
    for c in kat["cases"]:
        n, mn, mx, lo_p, hi_p, schema = c["args"]
        size = B._capi.lib().bsq_synth_size(n, mn, mx)
        assert size == c["bytes"]
        buf = torch.zeros(size + 64, dtype=torch.uint8, device="cuda:0")
        w = gpu.synth_device(buf.data_ptr(), size, n, 0, n, mn, mx, lo_p, hi_p, B.parse_schema(schema))
        assert w == size
        host = buf[:size].cpu().numpy()
        assert oracle.sha256(host) == c["sha256"]
    # a slice of a large stream equals the same slice generated on the CPU
    n = 33659618
    part = oracle.synth(n, 150, 150, 2, 40, "illumina_1.8", first=n - 1000, count=1000)
    buf = torch.zeros(part.size, dtype=torch.uint8, device="cuda:0")
    gpu.synth_device(buf.data_ptr(), part.size, n, n - 1000, 1000, 150, 150, 2, 40, B.parse_schema("illumina_1.8"))
    assert np.array_equal(buf.cpu().numpy(), part)
    assert B._capi.lib().bsq_compute_num_reads_for_size(10 << 30, 150, 150) == 33659618
    gpu.close()


# ------------------------------------------------------------------ scale: size-independent properties


def _synth_on_device(B, gpu, n, mn, mx, schema="illumina_1.8"):
    import torch
    size = B._capi.lib().bsq_synth_size(n, mn, mx)
    buf = torch.empty(size + 256, dtype=torch.uint8, device="cuda:0")
    assert gpu.synth_device(buf.data_ptr(), size, n, 0, n, mn, mx, 2, 40, B.parse_schema(schema)) == size
    return buf, size


class _DevPtr:
    """Wraps an arena pointer as a torch tensor through the CUDA array interface."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def test_scale_fixed_length_soa_equals_strided_gather(B):
    """~1 GB of 150 bp records: the packed SoA must equal a strided gather of the input (records
    have a fixed size: header, 150 seq, '+', 150 qual), checked on the device."""
    import torch
    n = 3_000_000
    digits = len(str(n - 1))
    hdr, rec, idl = 7 + digits, 7 + digits + 304, 5 + digits
    gpu = B.GpuParser(True, True, B.parse_schema("sanger"), 4096)
    buf, size = _synth_on_device(B, gpu, n, 150, 150)
    assert size == n * rec
    res = gpu.parse_device(buf.data_ptr(), size, want=3)
    assert (res.n_records, res.n_bases, res.stop.code, res.n_batches) == (n, n * 150, 6, (n + 4095) // 4096)
    soa = gpu.soa_view()
    assert (soa.num_records, soa.seq_len, soa.total_id_bytes) == (n, n * 150, n * idl)
    seq = torch.as_tensor(_DevPtr(soa.sequence_buffer, n * 150, "|u1"), device="cuda:0")
    qual = torch.as_tensor(_DevPtr(soa.qual_buffer, n * 150, "|u1"), device="cuda:0")
    ids = torch.as_tensor(_DevPtr(soa.id_buffer, n * idl, "|u1"), device="cuda:0")
    ends = torch.as_tensor(_DevPtr(soa.ends, n, "<i8"), device="cuda:0")
    id_ends = torch.as_tensor(_DevPtr(soa.id_ends, n, "<i8"), device="cuda:0")
    recs = buf[:size].view(n, rec)
    assert torch.equal(seq.view(n, 150), recs[:, hdr:hdr + 150])
    assert torch.equal(qual.view(n, 150), recs[:, hdr + 153:hdr + 303])
    assert torch.equal(ids.view(n, idl), recs[:, 1:1 + idl])
    k = torch.arange(n, device="cuda:0")
    assert torch.equal(ends, (k % 4096 + 1) * 150)
    assert torch.equal(id_ends, (k % 4096 + 1) * idl)
    gpu.close()


def test_scale_two_windows(B, oracle):
    """> 2 GiB on the device: the pass is cut into two windows at a record boundary."""
    import torch
    n = 7_500_000  # 7.5M x 318 B = 2.39 GB
    gpu = B.GpuParser(False, False, B.parse_schema("illumina_1.8"), 4096)
    buf, size = _synth_on_device(B, gpu, n, 150, 150)
    res = gpu.parse_device(buf.data_ptr(), size, want=3)
    assert (res.n_records, res.n_bases, res.stop.code, res.n_windows) == (n, n * 150, 6, 2)
    # the last batch, through the C ABI, against the oracle on the same bytes
    lastb = int(res.n_batches) - 1
    first = lastb * 4096
    tail = buf[first * 318:size].cpu().numpy()
    views, bases, err = oracle.parse_all(tail)
    exp = oracle.build_batch(tail, views)
    got = gpu.batch_to_host(lastb)
    for g, e in zip(got, (exp[1], exp[2], exp[0], exp[4], exp[3])):
        assert np.array_equal(g, e)
    # offsets of the second window start right after the first window's last record
    v0, le0, sp0 = gpu.offsets_to_host(0)
    v1, le1, sp1 = gpu.offsets_to_host(1)
    assert int(v0.n_records) + int(v1.n_records) == n and int(v1.first_record) == int(v0.n_records)
    assert int(v1.stream_base) + int(le1[0]) + 1 == int(v0.n_records) * 318
    gpu.close()


def test_whole_batches_regions_and_the_host_batch_pipeline(B, oracle):
    """BSQ_WANT_WHOLE_BATCHES: a region that does not end the stream leaves its trailing partial batch unconsumed, so
    batches cut region by region are the batches of the whole stream; HostBatchPipeline (two parser handles alternating
    regions, D2H of one overlapping H2D + passes of the other) fills the same five arrays as one pass over everything."""
    from blazeseq_b200 import _capi as capi
    data = oracle.synth(30000, 60, 260, 2, 40, "sanger")
    views, bases, err = oracle.parse_all(data)
    m = 700
    exp = [oracle.build_batch(data, views[a:a + m]) for a in range(0, len(views), m)]   # (id, seq, qual, id_ends, ends)
    gpu = B.GpuParser(True, True, B.parse_schema("sanger"), m)
    pos = rec = 0
    region = 900_000
    while True:
        end = min(data.size, pos + region)
        last = end == data.size
        r = gpu.parse_host(data[pos:end], pos, rec, last, capi.WANT_BATCHES | capi.WANT_WHOLE_BATCHES)
        n = int(r.n_records)
        assert last or (n % m == 0 and n > 0 and r.stop.code == capi.OK)
        assert pos + int(r.bytes_consumed) == (int(views[rec + n]["header_start"]) if rec + n < len(views) else data.size)
        for b in range(int(r.n_batches)):
            got = gpu.batch_to_host(b)
            e = exp[rec // m + b]
            for g, x in zip(got, (e[1], e[2], e[0], e[4], e[3])):
                assert np.array_equal(g, x)
        pos += int(r.bytes_consumed)
        rec += n
        if last:
            assert r.stop.code == capi.EOF
            break
    assert rec == len(views)
    gpu.close()

    total = {k: sum(e[i].size for e in exp) for k, i in (("id", 0), ("seq", 1), ("qual", 2))}
    for region in (700_000, 2_000_000, 1 << 30):
        pipe = B.HostBatchPipeline(lambda: B.GpuParser(True, True, B.parse_schema("sanger"), m), region_bytes=region)
        seq = np.zeros(total["seq"], np.uint8); qual = np.zeros(total["qual"], np.uint8); idb = np.zeros(total["id"], np.uint8)
        ends = np.zeros(len(views), np.int64); id_ends = np.zeros(len(views), np.int64)
        for _ in range(2):     # the handles are reusable
            got = pipe.run(data, seq, qual, idb, ends, id_ends)
            assert got == (len(views), total["seq"], total["qual"], total["id"]) and pipe.stop.code == capi.EOF
            assert np.array_equal(seq, np.concatenate([e[1] for e in exp])) and np.array_equal(qual, np.concatenate([e[2] for e in exp]))
            assert np.array_equal(idb, np.concatenate([e[0] for e in exp]))
            assert np.array_equal(ends, np.concatenate([e[4] for e in exp])) and np.array_equal(id_ends, np.concatenate([e[3] for e in exp]))
        pipe.close()
    # an error stops the pipeline with the reference's context
    bad = data.copy()
    bad[int(views[20000]["seq_start"]) + 5] = 0x80
    pipe = B.HostBatchPipeline(lambda: B.GpuParser(True, True, B.parse_schema("sanger"), m), region_bytes=700_000)
    got = pipe.run(bad, None, None, None, None, None)
    _, _, oerr = oracle.parse_all(bad, oracle.config(True, True, "sanger"))
    assert got[0] == 20000 and pipe.stop.code == 4 and pipe.stop.message == oerr.message
    pipe.close()


def test_device_writer_round_trip(B, oracle):
    """bsq_write_records (FastqRecord.write over the device SoA, record.mojo:384-402): the text the device writes is the
    records' canonical four-line form -- equal to the input when the input is canonical, the '\r'-free / bare-'+' form
    otherwise -- its offsets are the running byte_len, and parsing it again gives the same batches
    (tests/fastq/test_fastq_integration.mojo round trips)."""
    import torch
    from blazeseq_b200 import _capi as capi
    m = 300
    data = oracle.synth(5000, 40, 260, 2, 40, "sanger")
    gpu = B.GpuParser(True, True, B.parse_schema("sanger"), m)
    r = gpu.parse_host(data, want=capi.WANT_BATCHES)
    assert r.n_records == 5000
    text, offs = gpu.write_records(want_offsets=True)
    assert np.array_equal(text, data)                                   # synthetic input is canonical FASTQ
    views, _, _ = oracle.parse_all(data)
    assert np.array_equal(offs[:-1].astype(np.int64), views["header_start"]) and int(offs[-1]) == data.size
    # a slice that starts inside a batch and ends inside another, into a caller's device buffer
    a, n = 257, 1234
    dev = torch.zeros(int(offs[a + n] - offs[a]) + 7, dtype=torch.uint8, device="cuda")
    t2, o2 = gpu.write_records(a, n, out_device_ptr=dev.data_ptr(), capacity=dev.numel(), want_offsets=True)
    assert np.array_equal(t2, data[int(offs[a]):int(offs[a + n])]) and np.array_equal(dev.cpu().numpy()[:t2.size], t2)
    assert np.array_equal(o2, offs[a:a + n + 1] - offs[a]) and not dev[t2.size:].any()
    with pytest.raises(Exception):
        gpu.write_records(a, n, out_device_ptr=dev.data_ptr(), capacity=10)      # too small a buffer
    b3 = B.DeviceFastqBatch(gpu, 3, gpu.batch_view(3))
    assert b3.write() == data[int(offs[3 * m]):int(offs[4 * m])].tobytes()

    # CRLF input with '+id' lines and padded ids: the writer emits what FastqRecord.write would
    recs = []
    for i in range(700):
        L = 30 + (i * 7) % 90
        seq = bytes(b"ACGT"[(i + j) % 4] for j in range(L))
        qual = bytes(33 + (i * 3 + j) % 40 for j in range(L))
        recs.append((b"r%d  desc %d" % (i, i), seq, qual))
    messy = b"".join(b"@" + i + b" \r\n" + s + b"\r\n+" + i + b"\r\n" + q + b"\r\n" for i, s, q in recs)
    arr = np.frombuffer(messy, np.uint8)
    gpu.close()
    gpu = B.GpuParser(False, False, B.parse_schema("sanger"), m)        # ('\r' in the quality line fails check_quality, Q6)
    r = gpu.parse_host(arr, want=capi.WANT_BATCHES)
    assert r.n_records == 700
    text, _ = gpu.write_records()
    ov, _, _ = oracle.parse_all(arr)
    ob = oracle.build_batch(arr, ov)
    ide = np.concatenate([[0], ob[3]]); ee = np.concatenate([[0], ob[4]])
    want = b"".join(b"@" + ob[0][ide[k]:ide[k + 1]].tobytes() + b"\n" + ob[1][ee[k]:ee[k + 1]].tobytes() + b"\n+\n" +
                    ob[2][ee[k]:ee[k + 1]].tobytes() + b"\n" for k in range(700))
    assert text.tobytes() == want
    # ... and the text parses back to the same records (ids stripped, '\r' kept in sequence / quality as the reference does)
    r2 = gpu.parse_host(text.copy(), want=capi.WANT_BATCHES)
    assert r2.n_records == 700
    for b in range(int(r2.n_batches)):
        got = gpu.batch_to_host(b)
        e = oracle.build_batch(arr, ov[b * m:(b + 1) * m])
        for g, x in zip(got, (e[1], e[2], e[0], e[4], e[3])):
            assert np.array_equal(g, x)
    gpu.close()


def test_cpp_runner_counts_and_errors(B, oracle, tmp_path, golden_dir):
    """examples/run_blazeseq.cpp over include/blazeseq_gpu.hpp: the "<records> <base_pairs>" line of the reference's benchmark
    runners (run_blazeseq.mojo / _batch / _gzip) in every mode, on plain / gzip / BGZF files, and the reference's error text
    after the records before the error."""
    import gzip
    import subprocess
    from blazeseq_b200 import bgzf
    from test_host_logic import _build_cpp_runner
    exe = _build_cpp_runner(str(tmp_path))
    data = oracle.synth(30000, 50, 250, 2, 40, "sanger")
    views, bases, err = oracle.parse_all(data)
    (tmp_path / "a.fastq").write_bytes(data.tobytes())
    (tmp_path / "a.fastq.gz").write_bytes(gzip.compress(data.tobytes(), 6))
    (tmp_path / "b.fastq.bgz").write_bytes(bgzf.compress(data.tobytes(), level=6))
    want = [str(len(views)), str(bases)]
    for name in ("a.fastq", "a.fastq.gz", "b.fastq.bgz"):
        for mode in ("views", "batches", "device_batches"):
            r = subprocess.run([exe, str(tmp_path / name), mode, "1000"], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0 and r.stdout.split() == want, (name, mode, r.stdout, r.stderr)
    r = subprocess.run([exe, os.path.join(golden_dir, "corpus", "example.fastq"), "views", "4096", "validate"], capture_output=True, text=True)
    ev, eb, _ = oracle.parse_all(np.fromfile(os.path.join(golden_dir, "corpus", "example.fastq"), np.uint8), oracle.config(True, True))
    assert r.returncode == 0 and r.stdout.split() == [str(len(ev)), str(eb)]
    # an error: the records before it are counted, then the reference's message
    bad = data.copy()
    bad[int(views[12345]["header_start"])] = ord("X")
    (tmp_path / "bad.fastq").write_bytes(bad.tobytes())
    bv, bb, berr = oracle.parse_all(bad)
    for mode in ("views", "batches"):
        r = subprocess.run([exe, str(tmp_path / "bad.fastq"), mode, "1000"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 2 and r.stdout.split() == [str(len(bv)), str(bb)], (mode, r.stdout)
        assert r.stderr.strip() == berr.message.decode().strip()


def test_shard_summaries_locate_record_starts(B, oracle):
    """Multi-GPU stitching: summaries of arbitrary byte shards (device) -> where each shard's first
    own record starts (host arithmetic) must match the oracle's record table."""
    import torch
    data = oracle.synth(40000, 75, 300, 2, 40, "illumina_1.8")
    views, bases, err = oracle.parse_all(data)
    starts = views["header_start"]
    dev = torch.from_numpy(data.copy()).cuda()
    gpu = B.GpuParser()
    rng = np.random.default_rng(3)
    for nshards in (2, 3, 8):
        cuts = np.sort(rng.integers(1, data.size - 1, nshards - 1))
        bounds = [0] + cuts.tolist() + [data.size]
        sums = [gpu.summarize_device(dev.data_ptr() + bounds[i], bounds[i + 1] - bounds[i]) for i in range(nshards)]
        st = B.shard_prefix(sums, [bounds[i + 1] - bounds[i] for i in range(nshards)])
        for i in range(nshards):
            first_own = int(np.searchsorted(starts, bounds[i]))   # first record starting at/after the cut
            assert st[i].first_record == first_own
            if first_own < len(starts) and starts[first_own] < bounds[i + 1]:
                assert bounds[i] + st[i].skip_bytes == starts[first_own]
            assert st[i].newline_rank == int((data[:bounds[i]] == 10).sum())
    gpu.close()


# ------------------------------------------------------------------ FastqParser streaming (several passes)


def _shard_worker(rank, world, port, data_bytes, cuts, halo, out_q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import blazeseq_b200 as B
    from blazeseq_b200 import sharding
    import gpu_util as U
    import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # both ranks drive cuda:0: gloo carries the 72 bytes
    data = np.frombuffer(data_bytes, np.uint8)
    views, bases, err = O.parse_all(data, O.config(buffer_growth_enabled=True))
    bounds = [0] + list(cuts) + [data.size]
    lo, hi = bounds[rank], bounds[rank + 1]
    end = min(data.size, hi + halo)
    shard = torch.from_numpy(data[lo:end].copy()).cuda()              # own bytes + halo
    gpu = B.GpuParser(batch_size=512, buffer_growth_enabled=True)
    plan = sharding.plan(dist, gpu.summarize_device(shard.data_ptr(), hi - lo), hi - lo)
    n = plan.end - plan.begin
    assert plan.end <= end - lo
    res = gpu.parse_device(shard.data_ptr() + plan.begin, n, lo + plan.begin, plan.first_record, rank == world - 1, 3)
    ok = res.bytes_consumed == n and res.stop.code == (O.EOF if rank == world - 1 else O.OK)
    k0, cnt = plan.first_record, int(res.n_records)
    g = U.gpu_offsets(gpu, res)                                        # absolute stream offsets (stream_base carries lo + begin)
    mine = views[k0:k0 + cnt]
    for name in U.NAMES5 + ("id_start", "id_len"):
        ok = ok and np.array_equal(g[name], mine[name])
    for b in range(int(res.n_batches)):
        seq, qual, idb, ends, id_ends = gpu.batch_to_host(b)
        oi, os_, oq, oie, oe = O.build_batch(data, mine[b * 512:(b + 1) * 512])
        ok = ok and all(np.array_equal(x, y) for x, y in ((seq, os_), (qual, oq), (idb, oi), (ends, oe), (id_ends, oie)))
    reads, total_bases = sharding.allreduce_counts(dist, cnt, int(res.n_bases))
    out_q.put((rank, k0, cnt, reads, total_bases, bool(ok), len(views), bases))
    gpu.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_parse_on_device(oracle, world):
    """SURVEY 8e on the device: one mixed-length stream cut at arbitrary byte offsets, one process per shard.  Every
    rank summarises its shard (bsq_summarize_device), the summaries are all-gathered, bsq_shard_prefix gives the cut
    points, the rank parses its own records + halo (bsq_parse_device) -- offsets and SoA batches equal the oracle's
    slice for that shard, and the all-reduced totals equal the whole stream's."""
    import socket
    import torch.multiprocessing as mp
    data = oracle.synth(60000, 75, 300, 2, 40, "illumina_1.8")
    rng = np.random.default_rng(world)
    cuts = sorted(int(x) for x in rng.integers(1, data.size - 1, world - 1))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, data.tobytes(), cuts, 1024, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    first = 0
    for rank, k0, cnt, reads, total_bases, ok, n_all, bases in results:
        assert ok and k0 == first and (reads, total_bases) == (n_all, bases), (rank, k0, first, cnt, ok)
        first += cnt
    assert first == results[0][6]


@pytest.mark.parametrize("region_bytes", [700, 4096, 100000])
@pytest.mark.parametrize("mutate", ["none", "crlf", "notail", "noise"])
def test_fastq_parser_streams_in_regions(B, oracle, region_bytes, mutate):
    """The host mirror reads the Reader in regions and carries the unconsumed tail (BufferedReader
    semantics); records, order and the final error must not depend on the region size."""
    rng = np.random.default_rng(region_bytes + len(mutate))
    data = _rand_stream(rng, 400, mutate, maxlen=120)
    cfg = oracle.config(True, True, "generic", buffer_growth_enabled=True)
    views, bases, err = oracle.parse_all(data, cfg)
    arr = np.frombuffer(data, np.uint8)
    p = B.FastqParser(B.MemoryReader(data), config=B.ParserConfig(check_ascii=True, check_quality=True,
                                                                  buffer_growth_enabled=True),
                      region_bytes=region_bytes)
    got = 0
    with pytest.raises(B.BlazeSeqError) as ei:
        while True:
            v = p.next_view()
            o = views[got]
            assert v.id() == bytes(arr[o["id_start"]:o["id_start"] + o["id_len"]])
            assert v.sequence() == bytes(arr[o["seq_start"]:o["seq_start"] + o["seq_len"]])
            assert v.quality() == bytes(arr[o["qual_start"]:o["qual_start"] + o["qual_len"]])
            got += 1
    assert got == len(views)
    assert str(ei.value) == err.text
    assert (ei.value.record_number, ei.value.line_number, ei.value.file_position) == \
        (err.record_number, err.line_number, err.file_position)


def test_device_batches_in_every_region_and_stale_views(B, oracle, tmp_path):
    """batches() stays on the device fast path after the first region (every batch of every region hands out
    its DeviceFastqBatch), also when the records of a trailing partial batch are carried into the next region and
    the carry is larger than the room in front of the pinned region buffer; a DeviceFastqBatch of an earlier pass
    refuses to be read once the parser has moved on."""
    data = oracle.synth(3000, 140, 160, 2, 40, "sanger")
    views, _, _ = oracle.parse_all(data, oracle.config())
    path = tmp_path / "r.fastq"
    path.write_bytes(data.tobytes())
    for reader, region in ((B.MemoryReader(data.tobytes()), 40000), (B.MemoryReader(data.tobytes()), 100000),
                           (B.FileReader(str(path)), 64 << 10), (B.FileReader(str(path)), 20000)):
        p = B.FastqParser(reader, batch_size=150, schema="sanger", region_bytes=region)   # 64 KiB: carry > 16 KiB
        seen, first_dev = 0, None
        for batch in p.batches(150):
            dev = batch.to_device()
            assert dev is not None and dev.valid(), seen                # device path in every region
            first_dev = first_dev or dev
            oi, os_, oq, oie, oe = oracle.build_batch(data, views[seen:seen + len(batch)])
            assert np.array_equal(batch._sequence_bytes, os_) and np.array_equal(batch._ends, oe)
            assert np.array_equal(batch._id_bytes, oi) and np.array_equal(batch._id_ends, oie)
            seen += len(batch)
        assert seen == 3000
        assert not first_dev.valid()
        with pytest.raises(B.BlazeSeqError, match="stale"):
            first_dev.copy_to_host()
        p.close()


@pytest.mark.parametrize("region_bytes", [5000, 1 << 30])
def test_fastq_parser_batches_across_regions(B, oracle, region_bytes):
    """batches(m) over a stream that needs several passes: every batch equals the oracle's FastqBatch."""
    data = oracle.synth(3000, 20, 90, 2, 40, "sanger")
    views, bases, err = oracle.parse_all(data)
    for m in (64, 1000):
        p = B.FastqParser(B.MemoryReader(data), batch_size=m, schema="sanger", region_bytes=region_bytes)
        first = 0
        for batch in p.batches():
            n = len(batch)
            oi, os_, oq, oie, oe = oracle.build_batch(data, views[first:first + n])
            assert np.array_equal(batch._ends, oe) and np.array_equal(batch._id_ends, oie)
            assert np.array_equal(batch._sequence_bytes, os_) and np.array_equal(batch._quality_bytes, oq)
            assert np.array_equal(batch._id_bytes, oi)
            assert n == min(m, len(views) - first)
            first += n
        assert first == len(views) and not p.has_more()


def test_fastq_parser_file_and_gzip_readers(B, oracle, tmp_path):
    import gzip
    data = oracle.synth(2000, 50, 150, 2, 40, "illumina_1.8").tobytes()
    (tmp_path / "a.fastq").write_bytes(data)
    with gzip.open(tmp_path / "a.fastq.gz", "wb") as f:
        f.write(data)
    views, bases, err = oracle.parse_all(data)
    for path in ("a.fastq", "a.fastq.gz"):
        p = B.parser(str(tmp_path / path), "illumina_1.8")
        n = nb = 0
        for rec in p.records():
            n += 1
            nb += len(rec)
        assert (n, nb) == (len(views), bases)     # the "<records> <base_pairs>" cross-check of the reference
        p = B.parser(str(tmp_path / path), "illumina_1.8")
        assert sum(len(b) for b in p.batches(512)) == len(views)


# ------------------------------------------------------------------ native file / gzip pipeline


@pytest.mark.parametrize("gz", [False, True, "bgzf"])
@pytest.mark.parametrize("region_bytes", [64 << 10, 1 << 20, 1 << 26])
def test_stream_pipeline_regions(B, oracle, tmp_path, gz, region_bytes):
    """bsq_stream_*: reader thread -> pinned regions -> passes.  Every region's tables must be the
    oracle's slice, batches stay whole across regions, totals and the stop reason agree."""
    import gzip
    from blazeseq_b200 import _capi as capi
    data = oracle.synth(12000, 75, 300, 2, 40, "illumina_1.8")
    path = tmp_path / ("s.fastq.gz" if gz else "s.fastq")
    if gz == "bgzf":       # members of <= 64 KiB, inflated block-parallel by the reader's worker threads
        from blazeseq_b200 import bgzf
        path.write_bytes(bgzf.compress(data.tobytes(), level=1))
        assert gzip.decompress(path.read_bytes()) == data.tobytes()
    elif gz:
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(data.tobytes())
    else:
        path.write_bytes(data.tobytes())
    views, bases, err = oracle.parse_all(data)
    m = 512
    gpu = B.GpuParser(False, False, B.parse_schema("illumina_1.8"), m)
    st = gpu.stream_open(str(path), capi.SOURCE_AUTO, region_bytes)
    done = 0
    while True:
        res, region, off, first = gpu.stream_next(st, capi.WANT_OFFSETS | capi.WANT_BATCHES)
        n = int(res.n_records)
        assert first == done and off == (int(views[done]["header_start"]) if done < len(views) else data.size)
        assert np.array_equal(region[:int(res.bytes_consumed)], data[off:off + int(res.bytes_consumed)])
        if res.stop.code == capi.OK:
            assert n % m == 0 or n < m            # batches are not split by a region boundary
        for b in range(int(res.n_batches)):
            got = gpu.batch_to_host(b)
            exp = oracle.build_batch(data, views[done + b * m: done + min((b + 1) * m, n)])
            for g, e in zip(got, (exp[1], exp[2], exp[0], exp[4], exp[3])):
                assert np.array_equal(g, e)
        done += n
        if res.stop.code != capi.OK:
            assert res.stop.code == capi.EOF
            break
    assert done == len(views)
    stats = gpu.stream_stats(st)
    assert stats.bytes_read == data.size and stats.regions >= 1
    gpu.stream_close(st)
    gpu.close()


def test_bgzf_parallel_inflate(B, oracle, tmp_path, golden_dir):
    """BGZF input: the reference's own .bgz fixtures and a synthetic file, 1 / 3 / all inflate threads; a
    corrupted member is reported, not parsed."""
    from blazeseq_b200 import _capi as capi, bgzf
    for f in ("example.fastq.bgz", "example_dos.fastq.bgz"):
        plain = open(os.path.join(golden_dir, "corpus", f[:-4]), "rb").read()
        exp = [(r["id_len"], r["seq_len"]) for r in oracle.parse_all(plain)[0]]
        p = B.FastqParser(B.RapidgzipReader(os.path.join(golden_dir, "corpus", f), 2), "generic")
        assert [(len(r.id), len(r.sequence)) for r in p.records()] == [(int(a), int(b)) for a, b in exp]
    data = oracle.synth(30000, 50, 250, 2, 40, "sanger")
    views, bases, err = oracle.parse_all(data)
    path = tmp_path / "p.fastq.gz"
    blob = bgzf.compress(data.tobytes(), level=1)
    path.write_bytes(blob)
    for threads, host_inflate in ((1, True), (3, True), (0, True), (0, False)):
        p = B.FastqParser(B.RapidgzipReader(str(path), threads), "sanger", region_bytes=1 << 20, host_inflate=host_inflate,
                          config=B.ParserConfig(check_ascii=True, check_quality=True))
        n = nb = 0
        for batch in p.batches(1000):
            n += len(batch)
            nb += batch.seq_len()
        assert (n, nb) == (len(views), bases)
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x55
    path.write_bytes(bytes(bad))
    for host_inflate in (True, False):
        gpu = B.GpuParser(False, False, B.parse_schema("sanger"), 512, host_inflate=host_inflate)
        st = gpu.stream_open(str(path), capi.SOURCE_AUTO, 1 << 20)
        with pytest.raises(Exception):
            while True:
                res, region, off, first = gpu.stream_next(st, capi.WANT_OFFSETS)
                if res.stop.code != capi.OK:
                    break
        gpu.stream_close(st)
        gpu.close()


def test_plain_gzip_parallel_decode_feeds_the_passes(B, oracle, tmp_path):
    """An ordinary (non-BGZF) .gz through the stream pipeline: `parallelism` host threads decode it speculatively
    (bsq_pgzip.h; 1 = zlib's gzread), regions / carries / batches as for any other source; a damaged stream is
    reported as a read failure, never parsed."""
    import gzip
    from blazeseq_b200 import _capi as capi
    data = oracle.synth(40000, 100, 250, 2, 40, "sanger")          # ~7 MB compressed: several 2 MiB chunks
    views, bases, err = oracle.parse_all(data)
    path = tmp_path / "q.fastq.gz"
    blob = gzip.compress(data.tobytes(), 6)
    path.write_bytes(blob)
    for threads in (1, 2, 0):
        p = B.FastqParser(B.RapidgzipReader(str(path), threads), "sanger", region_bytes=3 << 20,
                          config=B.ParserConfig(check_ascii=True, check_quality=True))
        n = nb = 0
        first_ids = []
        for batch in p.batches(1000):
            if n == 0:
                first_ids = [batch.get_record(i).id for i in range(3)]
            n += len(batch)
            nb += batch.seq_len()
        assert (n, nb) == (len(views), bases)
        assert first_ids == [bytes(data[int(v["id_start"]):int(v["id_start"]) + int(v["id_len"])]).decode() for v in views[:3]]
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x55
    path.write_bytes(bytes(bad))
    # (a streaming decoder hands bytes on before the member's CRC-32 is known, as gzread does: the damage surfaces either
    #  as a read failure or, earlier, as a parse error on the garbage -- never as a clean end of the stream)
    gpu = B.GpuParser(False, False, B.parse_schema("sanger"), 512)
    st = gpu.stream_open(str(path), capi.SOURCE_AUTO, 1 << 20)
    outcome = None
    try:
        while True:
            res, region, off, first = gpu.stream_next(st, capi.WANT_OFFSETS)
            if res.stop.code != capi.OK:
                outcome = res.stop.code
                break
    except Exception:
        outcome = "raised"
    assert outcome == "raised" or outcome not in (capi.OK, capi.EOF), outcome
    gpu.stream_close(st)
    gpu.close()


def test_device_inflate_is_bit_exact_zlib(B, tmp_path, golden_dir):
    """k_inflate_members against zlib on every kind of DEFLATE block: the reference's own .bgz fixtures, text at
    compression levels 1 / 6 / 9, stored blocks (level 0 and incompressible bytes), fixed-Huffman blocks (tiny
    members), long runs (overlapping matches), a byte alphabet that forces code lengths beyond the 10-bit lookup,
    an empty file; a payload bit flip and a wrong CRC are both reported as BSQ_E_IO."""
    import struct
    import zlib
    from blazeseq_b200 import _capi as capi, bgzf
    rng = np.random.default_rng(11)

    def inflate_on_device(blob: bytes):
        path = tmp_path / "x.bgz"
        path.write_bytes(blob)
        gpu = B.GpuParser(batch_size=64)
        st = gpu.stream_open(str(path), capi.SOURCE_GZIP, 256 << 20)
        try:
            res, region, off, first = gpu.stream_next(st, capi.WANT_OFFSETS)
            return bytes(region)
        finally:
            gpu.stream_close(st)
            gpu.close()

    def member(chunk: bytes, level: int, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        d = c.compress(chunk) + c.flush()
        assert len(d) + 26 <= 0x10000
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(d) + 25) + d +
                struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    for f in ("example.fastq.bgz", "example_dos.fastq.bgz"):
        blob = open(os.path.join(golden_dir, "corpus", f), "rb").read()
        d = zlib.decompressobj(31)
        exp = b""
        rest = blob
        while rest:                                  # every gzip member of the file
            d = zlib.decompressobj(31)
            exp += d.decompress(rest)
            rest = d.unused_data
        assert inflate_on_device(blob) == exp, f
    text = b"".join(b"@read_%07d some/description here\nACGTTGCANNACGT%s\n+\nIIIIHHHHFFFF%s\n" % (
        i, bytes(rng.choice(list(b"ACGT"), 80).astype(np.uint8)), bytes(rng.integers(35, 75, 80).astype(np.uint8)))
        for i in range(40000))
    skew = bytes(np.minimum(rng.geometric(0.02, 600000), 255).astype(np.uint8))     # ~250 symbols, very uneven: long codes
    cases = {
        "text level 1": (text, 1), "text level 6": (text, 6), "text level 9": (text, 9), "stored (level 0)": (text[:500000], 0),
        "incompressible": (rng.integers(0, 256, 700001, dtype=np.uint8).tobytes(), 6),
        "runs": (b"A" * 300000 + b"ab" * 100000 + b"xyz" * 70000 + bytes(200000), 6),
        "skewed alphabet": (skew, 9),
        "huffman only": (text[:400000], (6, zlib.Z_HUFFMAN_ONLY)), "fixed codes": (text[:300000], (6, zlib.Z_FIXED)),
    }
    for name, (data, level) in cases.items():
        strategy = zlib.Z_DEFAULT_STRATEGY
        if isinstance(level, tuple):
            level, strategy = level
        step = 0xFF00 if level else 60000
        blob = b"".join(member(data[i:i + step], level, strategy) for i in range(0, len(data), step)) + eof
        got = inflate_on_device(blob)
        assert got == data, name
    # tiny members (fixed Huffman), members of every size around the 32-lane slicing of the CRC kernel, empty input
    tiny = [text[i * 97:i * 97 + n] for i, n in enumerate(list(range(1, 70)) + [255, 256, 257, 1023, 1024, 1025, 4097])]
    assert inflate_on_device(b"".join(member(t, 6) for t in tiny) + eof) == b"".join(tiny)
    assert inflate_on_device(eof) == b""
    # damage: a flipped payload bit (caught by the decoder or by the CRC), a wrong CRC on intact data
    good = b"".join(member(text[i:i + 0xFF00], 6) for i in range(0, 400000, 0xFF00)) + eof
    flipped = bytearray(good); flipped[20000] ^= 0x10
    wrong_crc = bytearray(good); first_len = struct.unpack_from("<H", good, 16)[0] + 1; wrong_crc[first_len - 8] ^= 1
    for blob in (bytes(flipped), bytes(wrong_crc)):
        with pytest.raises(capi.BsqLibraryError):
            inflate_on_device(blob)


@pytest.mark.parametrize("src", ["example", "synthetic"])
def test_writer_roundtrips_like_the_reference_integration_tests(B, oracle, golden_dir, tmp_path, src):
    """tests/fastq/test_fastq_integration.mojo:143-268: parse -> write (plain / gzip / BGZF) -> parse again ->
    the records compare equal, for example.fastq and for a generated file; on the canonical generated file the
    written bytes are the input bytes."""
    import gzip
    from blazeseq_b200 import bgzf
    if src == "example":
        data = open(os.path.join(golden_dir, "corpus", "example.fastq"), "rb").read()
    else:
        data = oracle.synth(20000, 5, 300, 2, 40, "sanger").tobytes()
    (tmp_path / "in.fastq").write_bytes(data)
    original = list(B.parser(str(tmp_path / "in.fastq"), "sanger").records())
    assert len(original) == len(oracle.parse_all(data)[0])
    written = b"".join(r.write() for r in original)
    if src == "synthetic":
        assert written == data
    (tmp_path / "out.fastq").write_bytes(written)
    with gzip.open(tmp_path / "out.fastq.gz", "wb", compresslevel=1) as f:
        f.write(written)
    (tmp_path / "out.fastq.bgz").write_bytes(bgzf.compress(written, level=1))
    for name in ("out.fastq", "out.fastq.gz", "out.fastq.bgz"):
        path = str(tmp_path / name)
        reread = list((B.FastqGZParser(B.RapidgzipReader(path, 3), "sanger") if name.endswith("z")
                       else B.parser(path, "sanger")).records())
        assert reread == original, name
    # gzip -> plain -> gzip (test_gzip_roundtrip / test_gzip_to_gzip_roundtrip)
    from_gz = list(B.parser(str(tmp_path / "out.fastq.gz"), "sanger").records())
    again = b"".join(r.write() for r in from_gz)
    assert again == written
    batches = list(B.parser(str(tmp_path / "out.fastq.gz"), "sanger").batches(777))
    assert [r for b in batches for r in b.to_records()] == original


def test_native_and_python_io_paths_agree(B, oracle, tmp_path):
    data = oracle.synth(5000, 30, 120, 2, 40, "sanger").tobytes() + b"@tail\nACGT\n+\nIIII"   # no final newline
    (tmp_path / "t.fastq").write_bytes(data)
    out = []
    for native in (True, False):
        p = B.FastqParser(B.FileReader(tmp_path / "t.fastq"), "sanger", region_bytes=50000, native_io=native,
                          config=B.ParserConfig(check_ascii=True, check_quality=True, buffer_growth_enabled=True))
        recs = [(r.id, r.sequence, r.quality) for r in p.records()]
        out.append(recs)
        assert not p.has_more()
    assert out[0] == out[1] and len(out[0]) == 5001 and out[0][-1] == ("tail", "ACGT", "IIII")
