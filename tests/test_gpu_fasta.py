"""GPU parity of the FASTA path: bsq_fasta_* through the C ABI and the FastaParser mirror against the
oracle (ora_fasta_parse, a restatement of blazeseq/fasta/parser.mojo:60-200) on the reference's own literal streams,
its Biopython corpus and adversarial random streams.  Bit-exact: ids, sequences, record count, stop code, message
text, record / line number and file position."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from fasta_cases import CASES, CORPUS  # noqa: E402


@pytest.fixture(scope="module")
def B():
    import blazeseq_b200
    return blazeseq_b200


def check(B, oracle, data: bytes, check_ascii=False, gpu=None):
    arr = np.frombuffer(data, np.uint8)
    ids, seqs, err = oracle.fasta_parse(arr, check_ascii)
    own = gpu is None
    if own:
        gpu = B.GpuParser(check_ascii=check_ascii)
    res = gpu.fasta_parse_host(np.ascontiguousarray(arr))
    assert res.n_records == len(ids), (res.n_records, len(ids), res.stop.text, err.message)
    assert res.stop.code == err.code and res.stop.message == err.message, (res.stop.message, err.message)
    assert (res.stop.record_number, res.stop.line_number, res.stop.file_position) == (
        err.record_number, err.line_number, err.file_position)
    seq, ss, idb, ist = gpu.fasta_to_host()
    for i in range(len(ids)):
        assert idb[int(ist[i]):int(ist[i + 1])].tobytes() == ids[i], i
        assert seq[int(ss[i]):int(ss[i + 1])].tobytes() == seqs[i], i
    assert res.n_bases == sum(len(s) for s in seqs)
    if own:
        gpu.close()
    return res


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_literals(B, oracle, case):
    cite, data, check_ascii, recs, sub = case
    res = check(B, oracle, data, check_ascii)
    assert sub in res.stop.text
    p = B.FastaParser(B.MemoryReader(data), B.FastaParserConfig(check_ascii))
    for exp in recs:
        r = p.next_record()
        assert (r.id(), r.sequence()) == exp
    with pytest.raises(B.BlazeSeqError) as ei:
        p.next_record()
    assert sub in str(ei.value)


def test_biopython_corpus(B, oracle, golden_dir):
    for name in sorted(os.listdir(os.path.join(golden_dir, "fasta_corpus"))):
        if name == "README.md":
            continue
        data = open(os.path.join(golden_dir, "fasta_corpus", name), "rb").read()
        check(B, oracle, data)
        check(B, oracle, data, True)
    for name, count, checks in CORPUS:
        recs = list(B.FastaParser(B.FileReader(os.path.join(golden_dir, "fasta_corpus", name))))
        assert (len(recs) >= 1) if count is None else (len(recs) == count)
        for i, id_sub, seq_sub in checks:
            assert id_sub in recs[i].id() and (seq_sub is None or seq_sub in recs[i].sequence())


def _rand_fasta(rng, nrec, mutate):
    out = []
    for i in range(nrec):
        ident = b"seq%d some description %d" % (i, int(rng.integers(0, 1000)))
        L = int(rng.integers(1, 400))
        width = int(rng.choice([1, 7, 60, 61, 80, 1000]))
        seq = bytes(rng.choice(list(b"ACGTNacgt-*"), L).astype(np.uint8))
        nl = b"\r\n" if mutate == "crlf" else b"\n"
        lines = [seq[k:k + width] for k in range(0, L, width)]
        if mutate == "blanks":
            lines = [(b" " * int(rng.integers(0, 3))) + ln + (b"\t" * int(rng.integers(0, 2))) for ln in lines]
            ident = b"  " + ident + b" \t"
            if rng.random() < 0.3:
                lines.insert(int(rng.integers(0, len(lines) + 1)), b"")
        rec = b">" + ident + nl + nl.join(lines) + nl
        if mutate == "blanks" and rng.random() < 0.2:
            rec = nl + b"   " + nl + rec
        out.append(rec)
    data = b"".join(out)
    if mutate == "notail":
        data = data.rstrip(b"\r\n")
    if mutate == "empty" and nrec > 3:
        k = int(rng.integers(1, nrec))
        out[k] = b">empty_record" + (b"\n" if rng.random() < 0.5 else b"\n\n  \n")
        data = b"".join(out)
    if mutate == "junk":
        data = b"\n \nACGT not a header\n" + data
    if mutate == "hi" and len(data) > 10:
        b = bytearray(data)
        for _ in range(3):
            b[int(rng.integers(0, len(b)))] |= 0x80
        data = bytes(b)
    return data


@pytest.mark.parametrize("mutate", ["none", "crlf", "blanks", "notail", "empty", "junk", "hi"])
def test_random_fasta_streams(B, oracle, mutate):
    rng = np.random.default_rng(abs(hash("fa" + mutate)) % 2**32)
    for ascii_on in (False, True):
        gpu = B.GpuParser(check_ascii=ascii_on)
        for trial in range(6):
            data = _rand_fasta(rng, int(rng.integers(1, 2500 if trial == 5 else 300)), mutate)
            check(B, oracle, data, ascii_on, gpu=gpu)
        gpu.close()


def test_long_sequences_and_many_lines(B, oracle):
    rng = np.random.default_rng(7)
    big = bytes(rng.choice(list(b"ACGT"), 3_000_000).astype(np.uint8))
    lines = b"\n".join(big[k:k + 60] for k in range(0, len(big), 60))
    check(B, oracle, b">chr1 test\n" + lines + b"\n>chr2\n" + big[:100000] + b"\n>chr3\nAC\nGT")
    check(B, oracle, b">a\n" + b"\n" * 70000 + b"ACGT\n" + b"\n" * 1000)          # more newlines than a tile's list holds
