"""Pins the oracle (oracle/bsq_oracle.c) against the reference's own tests, fixtures and docs.

CPU-only.  Sources of truth (all committed under tests/golden/ by make_golden.py):
  * literal streams of tests/fastq/test_parser.mojo and tests/test_error_context.mojo (ref_cases.py)
  * the 70-file corpus with the README "Current error" table and the accept-sets of
    tests/fastq/test_fastq_parser_correctness.mojo
  * tests/test_python_bindings.py ids; tests/fastq/test_record_batch.mojo SoA layout
  * generator KATs
"""
import json
import os

import numpy as np
import pytest

from ref_cases import BATCH_CASES, CASES, EXAMPLE_IDS


def _cfg(O, kw):
    return O.config(check_ascii=kw.get("check_ascii", False),
                    check_quality=kw.get("check_quality", False),
                    schema_name=kw.get("schema", "generic"),
                    buffer_capacity=kw.get("buffer_capacity"),
                    buffer_growth_enabled=kw.get("buffer_growth_enabled", False),
                    buffer_max_capacity=kw.get("buffer_max_capacity"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("api", ["next_view", "next_record"])
def test_literal_streams_streaming_model(oracle, case, api):
    name, cite, data, kw, records, err_sub = case
    p = oracle.StreamParser(data, _cfg(oracle, kw))
    for exp in records:
        rc, v, e = getattr(p, api)()
        assert rc == oracle.OK, (cite, e.text)
        assert p.fields(v) == exp, cite
    rc, v, e = getattr(p, api)()
    assert rc != oracle.OK, cite
    assert err_sub in e.text, (cite, e.text)


@pytest.mark.parametrize("case", [c for c in CASES if "buffer_capacity" not in c[3] or
                                  c[3]["buffer_capacity"] >= 256], ids=lambda c: c[0])
def test_literal_streams_canonical(oracle, case):
    """ora_parse_all (whole-buffer semantics) agrees wherever the buffer holds the stream."""
    name, cite, data, kw, records, err_sub = case
    views, bases, e = oracle.parse_all(data, _cfg(oracle, kw))
    a = np.frombuffer(data, np.uint8)
    got = [(bytes(a[v["id_start"]:v["id_start"] + v["id_len"]]),
            bytes(a[v["seq_start"]:v["seq_start"] + v["seq_len"]]),
            bytes(a[v["qual_start"]:v["qual_start"] + v["qual_len"]])) for v in views]
    assert got == records, cite
    assert err_sub in e.text, (cite, e.text)


@pytest.mark.parametrize("cite,data,bs,sizes", BATCH_CASES, ids=[c[0] for c in BATCH_CASES])
def test_batches(oracle, cite, data, bs, sizes):
    p = oracle.StreamParser(data, oracle.config())
    got = []
    while p.has_more():  # _FastqParserBatchIter, parser.mojo:719-735
        rc, views, e = p.next_batch(bs)
        assert rc == oracle.OK
        if len(views) == 0:
            break
        got.append(len(views))
    assert got == sizes, cite
    assert not p.has_more()


def test_record_number_tracking_context(oracle):
    """tests/test_error_context.mojo:97-137 -> record 2, line 5, position 12."""
    from ref_cases import SECOND_BAD
    p = oracle.StreamParser(SECOND_BAD, oracle.config(True, True))
    rc, v, e = p.next_record()
    assert rc == oracle.OK and v.seq_len == 2
    rc, v, e = p.next_record()
    assert rc == oracle.ID_NO_AT
    assert (e.record_number, e.line_number, e.file_position) == (2, 5, 12)
    assert e.text == ("Sequence id line does not start with '@'\n  Record number: 2\n"
                      "  Line number: 5\n  File position: 12\n  Record snippet: r2\nGC\n+\n#$\n")


def test_validation_error_text(oracle):
    """parser.mojo:163-169,597-610 + errors.mojo:223-234."""
    from ref_cases import NON_ASCII
    p = oracle.StreamParser(NON_ASCII, oracle.config(True, False))
    rc, v, e = p.next_view()
    assert rc == oracle.ASCII_INVALID
    assert e.message == b"Non ASCII letters found\n  Record number: 1\n  Record snippet: r1\nA\xc8C"


def test_iterator_yields_zero_on_bad_first_record(oracle):
    """tests/test_error_context.mojo:81-94."""
    from ref_cases import INVALID_ID
    views, bases, e = oracle.parse_all(INVALID_ID, oracle.config(True, True))
    assert len(views) == 0 and e.code == oracle.ID_NO_AT


# ---------------------------------------------------------------------------- corpus


@pytest.fixture(scope="module")
def expect(golden_dir):
    return json.load(open(os.path.join(golden_dir, "corpus_expect.json")))


def _read(golden_dir, f):
    return open(os.path.join(golden_dir, "corpus", f), "rb").read()


# README "Current error" rows that the reference's own code contradicts:
# the first failing check on these files is the '+' check (utils.mojo:456) or there is no error.
README_STALE = {
    "error_double_seq.fastq": "Separator line does not start with '+'",
    "error_trunc_at_plus.fastq": "Separator line does not start with '+'",
    "error_trunc_at_seq.fastq": "Separator line does not start with '+'",
    "error_trunc_in_seq.fastq": "Separator line does not start with '+'",
    "error_trunc_in_title.fastq": "Separator line does not start with '+'",
    "zero_length.fastq": "EOF",
    # valid under the valid-file tests (validation OFF); the README column describes that run
    "example_dos.fastq": None,
}


def test_corpus_invalid_files(oracle, expect, golden_dir):
    """invalid_file_test_fun(_ref), test_fastq_parser_correctness.mojo:21-110: must stop with
    msg or one of the accept-set; and must equal the README 'Current error' unless stale."""
    accept = expect["accept_set"]
    n = 0
    for f, row in expect["files"].items():
        if not row["invalid_msg"]:
            continue
        n += 1
        data = _read(golden_dir, f)
        for runner in ("canonical", "stream"):
            if runner == "canonical":
                views, bases, e = oracle.parse_all(data, oracle.config(True, True))
                text = e.text
            else:
                p = oracle.StreamParser(data, oracle.config(True, True))
                while True:
                    rc, v, e = p.next_record()
                    if rc != oracle.OK:
                        break
                text = e.text
            assert row["invalid_msg"] in text or any(a in text for a in accept), (f, text)
            want = README_STALE.get(f, row["readme_current_error"])
            assert want in text, (f, runner, text)
    assert n == 30


def test_corpus_valid_files(oracle, expect, golden_dir):
    """valid_file_test_fun(_ref), test_fastq_parser_correctness.mojo:59-72: validation OFF, the
    file's schema; must reach EOF without an error."""
    n = 0
    for f, row in expect["files"].items():
        if not row["valid_schema"] or f in expect["multi_line_disabled"]:
            continue
        n += 1
        data = _read(golden_dir, f)
        views, bases, e = oracle.parse_all(data, oracle.config(False, False, row["valid_schema"]))
        assert e.code == oracle.EOF, (f, e.text)
        assert len(views) > 0
        # streaming model with the default 256 KiB buffer gives the same records
        p = oracle.StreamParser(data, oracle.config(False, False, row["valid_schema"]))
        for v in views:
            rc, sv, se = p.next_view()
            assert rc == oracle.OK
            assert all(getattr(sv, k) == v[k] for k in v.dtype.names), f
        rc, sv, se = p.next_view()
        assert rc == oracle.EOF
    assert n == 37


def test_example_ids(oracle, golden_dir):
    """tests/test_python_bindings.py:31-67."""
    data = _read(golden_dir, "example.fastq")
    p = oracle.StreamParser(data, oracle.config())
    rc, views, e = p.next_batch(2)
    assert [p.fields(v)[0] for v in views] == EXAMPLE_IDS[:2]
    rc, views, e = p.next_batch(10)
    assert [p.fields(v)[0] for v in views] == EXAMPLE_IDS[2:]
    assert b"CCCTTCTTGTCTTCAGCGTTTCTCC" in oracle.StreamParser(data).fields(
        oracle.parse_all(data)[0][0])[1]


def test_example_dos_crlf(oracle, golden_dir):
    """CRLF -> id loses '\\r' (strip), seq/qual keep it."""
    data = _read(golden_dir, "example_dos.fastq")
    views, bases, e = oracle.parse_all(data, oracle.config())
    assert len(views) == 3 and e.code == oracle.EOF
    p = oracle.StreamParser(data)
    i, s, q = p.fields(views[0])
    assert i == EXAMPLE_IDS[0] and s.endswith(b"\r") and q.endswith(b"\r")


def test_batch_layout(oracle):
    """tests/fastq/test_record_batch.mojo:26-38: _ends == [2, 4]; arrays of length 4."""
    data = b"@a\nAC\n+\n!!\n@b\nGT\n+\n!!\n"
    views, bases, e = oracle.parse_all(data)
    idb, sqb, qlb, ide, ends = oracle.build_batch(data, views)
    assert ends.tolist() == [2, 4] and ide.tolist() == [1, 2]
    assert bytes(sqb) == b"ACGT" and bytes(qlb) == b"!!!!" and bytes(idb) == b"ab"


# ---------------------------------------------------------------------------- tail rule (A.2)


@pytest.mark.parametrize("tail,code,sub", [
    (b"@r2\nAC\n+\nII", 0, None),                       # accepted, no structure check (Q1)
    (b"\n", 7, "at phase 1"), (b"\n\n", 7, "at phase 2"), (b"\n\n\n", 11, ""),
    (b"\n\n\n\n", 1, "does not start with '@'"),
    (b"@r2\nAC", 7, "at phase 1"), (b"@r2", 7, "at phase 0"), (b"@r2\nAC\n+", 7, "at phase 2"),
    (b"@r2\nAC\n+\n \t\r", 11, ""),
    (b"xx\nAC\nyy\nIIII", 0, None),                    # Q1: malformed but accepted at the tail
])
def test_tail_rule(oracle, tail, code, sub):
    data = b"@r1\nAC\n+\nII\n" + tail
    views, bases, e = oracle.parse_all(data, oracle.config())
    p = oracle.StreamParser(data, oracle.config())
    n = 0
    while True:
        rc, v, se = p.next_view()
        if rc != oracle.OK:
            break
        n += 1
    assert n == len(views)
    assert se.code == e.code and se.text == e.text
    if code == 0:
        assert len(views) == 2 and e.code == oracle.EOF
    else:
        assert len(views) == 1 and e.code == code and sub in e.text


def test_q2_single_record_without_newline(oracle):
    """first record of the buffer incomplete, growth off."""
    for fn in ("canonical", "stream"):
        data = b"@a\nAC\n+\n!!"
        if fn == "canonical":
            views, bases, e = oracle.parse_all(data, oracle.config())
            assert len(views) == 0
        else:
            rc, v, e = oracle.StreamParser(data, oracle.config()).next_view()
        assert e.code == oracle.BUFFER_EXCEEDED
        assert e.text == ("FASTQ record exceeds buffer capacity (262144 bytes). Enable buffer "
                          "growth or increase buffer_capacity.")
    views, bases, e = oracle.parse_all(data, oracle.config(buffer_growth_enabled=True))
    assert len(views) == 1 and e.code == oracle.EOF


def test_buffer_capacity_invisible_when_input_ends_with_newline(oracle):
    """capacities 16..199 over a 10-record stream give identical records."""
    data = oracle.synth(10, 3, 9, 2, 40, "sanger").tobytes()
    ref, _, _ = oracle.parse_all(data)
    longest = max(int(v["record_end"] - v["header_start"]) + 1 for v in ref)
    for cap in range(longest + 1, 200):
        p = oracle.StreamParser(data, oracle.config(buffer_capacity=cap))
        for v in ref:
            rc, sv, e = p.next_view()
            assert rc == oracle.OK, (cap, e.text)
            assert all(getattr(sv, k) == v[k] for k in v.dtype.names)
        assert p.next_view()[0] == oracle.EOF


def test_quality_compat_q5(oracle):
    """`~` (UPPER) accepted by intent; rejected in the SIMD body when emulated."""
    data = b"@r\n" + b"A" * 40 + b"\n+\n" + b"~" + b"I" * 39 + b"\n"
    assert oracle.parse_all(data, oracle.config(False, True, "sanger"))[2].code == oracle.EOF
    assert oracle.parse_all(data, oracle.config(False, True, "sanger", compat_simd_width=32))[
        2].code == oracle.QUALITY_OUT_OF_RANGE
    tail = b"@r\n" + b"A" * 40 + b"\n+\n" + b"I" * 39 + b"~" + b"\n"
    assert oracle.parse_all(tail, oracle.config(False, True, "sanger", compat_simd_width=32))[
        2].code == oracle.EOF


def test_generic_rejects_space_and_high_bit(oracle):
    """tests/fastq/test_fastq_record.mojo:133-176."""
    sp = b"@r\nACGT\n+\nII I\n"
    assert oracle.parse_all(sp, oracle.config(False, True))[2].code == oracle.QUALITY_OUT_OF_RANGE
    hb = b"@r\nACGT\n+\nII\x80I\n"
    assert oracle.parse_all(hb, oracle.config(True, False))[2].code == oracle.ASCII_INVALID


# ---------------------------------------------------------------------------- generator


def test_generator_kats(oracle, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "synthetic_kat.json")))
    for c in kat["cases"]:
        b = oracle.synth(*c["args"])
        assert b.size == c["bytes"] and oracle.sha256(b) == c["sha256"]
        import hashlib
        assert hashlib.sha256(b.tobytes()).hexdigest() == c["sha256"]
    for c in kat["compute_num_reads_for_size"]:
        assert oracle.compute_num_reads_for_size(*c["args"]) == c["reads"]
    # values SURVEY.md App. B.3 derived with an independent model of utils.mojo
    assert kat["cases"][0]["sha256"] == \
        "1901f93b73164d031ff71d63e406e6b70d69240b9b393480251774b2b80cb563"
    assert oracle.compute_num_reads_for_size(10 << 30, 150, 150) == 33659618
    assert oracle.compute_num_reads_for_size(3 << 30, 100, 100) == 14708792  # "14.7M reads" plots


def test_generator_parses_and_lengths(oracle):
    """tests/fastq/test_parser.mojo:228-258: 20 reads, len in [5,12], batches(8)."""
    b = oracle.synth(20, 5, 12, 2, 25, "generic")
    assert bytes(b[:38]) == b"@read_00\nGCATCGTGCTGA\n+\n7452141-')##\n@"
    views, bases, e = oracle.parse_all(b, oracle.config(True, True))
    assert len(views) == 20 and e.code == oracle.EOF
    assert views["seq_len"].min() >= 5 and views["seq_len"].max() <= 12


def test_generator_slices_concatenate(oracle):
    whole = oracle.synth(1000, 75, 300, 2, 40, "illumina_1.8")
    parts = [oracle.synth(1000, 75, 300, 2, 40, "illumina_1.8", first=f, count=250)
             for f in (0, 250, 500, 750)]
    assert np.array_equal(np.concatenate(parts), whole)


def test_mt_baseline_matches(oracle):
    b = oracle.synth(5000, 75, 300, 2, 40, "illumina_1.8")
    views, bases, e = oracle.parse_all(b, oracle.config(True, True))
    for threads in (1, 2, 3, 8):
        for mode in (0, 1):
            n, nb, code = oracle.baseline_mt(b, oracle.config(True, True), mode, 4096, threads)
            assert (n, nb, code) == (len(views), bases, 0)


def test_record_limit_of_the_canonical_parse_equals_the_streaming_model(oracle):
    """parser.mojo:484-503.  ora_parse_all reports a record longer than buffer_capacity (growth off) or
    buffer_max_capacity (growth on) as BUFFER_EXCEEDED / BUFFER_AT_MAX.  For streams whose records all end in a
    newline the BufferedReader model (ora_open / ora_next_view) must give the same records and the same stop."""
    rng = np.random.default_rng(1)

    def rec(i, L):
        return b"@r%d\n" % i + bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)) + b"\n+\n" + b"I" * L + b"\n"
    hits = 0
    for trial in range(300):
        cap = int(rng.integers(40, 300))
        growth = bool(trial % 2)
        mx = int(rng.integers(cap, 600))
        data = b"".join(rec(i, int(rng.integers(1, 160))) for i in range(int(rng.integers(1, 30))))
        cfg = oracle.config(False, False, "generic", buffer_capacity=cap, buffer_growth_enabled=growth,
                            buffer_max_capacity=mx)
        views, bases, err = oracle.parse_all(data, cfg)
        p = oracle.StreamParser(data, cfg)
        n = 0
        while True:
            rc, v, se = p.next_view()
            if rc != oracle.OK:
                break
            assert (v.header_start, v.record_end) == (int(views[n]["header_start"]), int(views[n]["record_end"]))
            n += 1
        assert n == len(views), (trial, cap, growth, mx)
        assert (se.code, se.message) == (err.code, err.message), (trial, cap, growth, mx, se.message, err.message)
        hits += err.code in (8, 9)
    assert hits > 50   # the limit was exercised


def test_fasta_oracle_against_the_reference_literals_and_corpus(oracle, golden_dir):
    """ora_fasta_parse (fasta/parser.mojo:60-200) against every literal stream of tests/fasta/test_fasta_parser.mojo and
    the Biopython files checked by tests/fasta/test_fasta_parser_correctness.mojo."""
    from fasta_cases import CASES, CORPUS
    for cite, data, check_ascii, recs, sub in CASES:
        ids, seqs, err = oracle.fasta_parse(data, check_ascii)
        assert list(zip(ids, seqs)) == recs, cite
        assert sub in err.message.decode("latin-1"), (cite, err.message)
    for name, count, checks in CORPUS:
        data = open(os.path.join(golden_dir, "fasta_corpus", name), "rb").read()
        ids, seqs, err = oracle.fasta_parse(data)
        assert (len(ids) >= 1) if count is None else (len(ids) == count), name
        assert err.code == oracle.EOF
        for i, id_sub, seq_sub in checks:
            assert id_sub in ids[i] and (seq_sub is None or seq_sub in seqs[i]), (name, i)
        assert all(b"\n" not in s and b"\r" not in s for s in seqs)
