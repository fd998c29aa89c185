"""Literal streams and expectations transcribed from the reference's OWN tests.

Each case: (name, citation, data, cfg kwargs, expected records [(id, seq, qual)],
expected terminal-error substring).  `cfg` keys follow ParserConfig (parser.mojo:33-74) plus
"schema".  The terminal error is what the next call after the listed records raises ("EOF" for a
clean end).  Used by the oracle tests (CPU) and by the GPU parity tests, which run the same
streams through the C-ABI.
"""

NON_ASCII = b"@r1\nA\xc8C\n+\n!!!\n"  # tests/fastq/test_parser.mojo:23-39
LONG20 = b"@id\n" + b"A" * 20 + b"\n+\n" + b"!" * 20 + b"\n"  # test_parser.mojo:514-524
INVALID_ID = b"r1\nATCG\n+\n!@#$\n"  # tests/test_error_context.mojo:10-28
MISMATCH = b"@r1\nATCG\n+\n!@#\n"  # tests/test_error_context.mojo:31-53
SECOND_BAD = b"@r1\nAT\n+\n!@\nr2\nGC\n+\n#$\n"  # tests/test_error_context.mojo:97-126

CASES = [
    ("for_loop_two_records", "tests/fastq/test_parser.mojo:42-67",
     b"@r1\nACGT\n+\n!!!!\n@r2\nTGCA\n+\n####\n", dict(schema="generic"),
     [(b"r1", b"ACGT", b"!!!!"), (b"r2", b"TGCA", b"####")], "EOF"),
    ("single_record_then_eof", "tests/fastq/test_parser.mojo:70-84",
     b"@r1\nACGT\n+\n!!!!\n", dict(schema="generic"), [(b"r1", b"ACGT", b"!!!!")], "EOF"),
    ("ascii_enabled", "tests/fastq/test_parser.mojo:87-98",
     NON_ASCII, dict(check_ascii=True, check_quality=False), [], "Non ASCII letters found"),
    ("ascii_disabled", "tests/fastq/test_parser.mojo:101-114",
     NON_ASCII, dict(check_ascii=False, check_quality=False), [(b"r1", b"A\xc8C", b"!!!")], "EOF"),
    ("batch_content", "tests/fastq/test_parser.mojo:163-177",
     b"@seq1\nACGT\n+\n!!!!\n", dict(schema="generic"), [(b"seq1", b"ACGT", b"!!!!")], "EOF"),
    ("empty_input", "tests/fastq/test_parser.mojo:180-192,450-458", b"", dict(), [], "EOF"),
    ("fast_path_cap256", "tests/fastq/test_parser.mojo:291-311",
     b"@r1\nACGT\n+\n!!!!\n@r2\nTGCA\n+\n!!!!\n", dict(buffer_capacity=256),
     [(b"r1", b"ACGT", b"!!!!"), (b"r2", b"TGCA", b"!!!!")], "EOF"),
    ("span_chunks_cap32", "tests/fastq/test_parser.mojo:327-341",
     b"@r1\nACGT\n+\n!!!!\n", dict(buffer_capacity=32), [(b"r1", b"ACGT", b"!!!!")], "EOF"),
    ("three_records_cap32", "tests/fastq/test_parser.mojo:364-385,492-511",
     b"@r1\nA\n+\n!\n@r2\nB\n+\n!\n@r3\nC\n+\n!\n", dict(buffer_capacity=32),
     [(b"r1", b"A", b"!"), (b"r2", b"B", b"!"), (b"r3", b"C", b"!")], "EOF"),
    ("two_records_cap32", "tests/fastq/test_parser.mojo:388-406",
     b"@a\nAC\n+\n!!\n@b\nTG\n+\n##\n", dict(buffer_capacity=32),
     [(b"a", b"AC", b"!!"), (b"b", b"TG", b"##")], "EOF"),
    ("invalid_header", "tests/fastq/test_parser.mojo:461-469",
     b"r1\nACGT\n+\n!!!!\n", dict(buffer_capacity=256), [],
     "Sequence id line does not start with '@'"),
    ("mismatched_len", "tests/fastq/test_parser.mojo:472-482",
     b"@r1\nACGT\n+\n!!!\n", dict(buffer_capacity=256), [],
     "Quality and sequence line do not match in length"),
    ("long_line_growth", "tests/fastq/test_parser.mojo:527-543",
     LONG20, dict(buffer_capacity=16, buffer_growth_enabled=True, buffer_max_capacity=256),
     [(b"id", b"A" * 20, b"!" * 20)], "EOF"),
    ("long_line_no_growth", "tests/fastq/test_parser.mojo:546-558",
     LONG20, dict(buffer_capacity=16, buffer_growth_enabled=False), [],
     "record exceeds buffer capacity"),
    ("ctx_invalid_id", "tests/test_error_context.mojo:56-66,140-149",
     INVALID_ID, dict(check_ascii=True, check_quality=True), [], "Record number"),
    ("ctx_invalid_id_line", "tests/test_error_context.mojo:140-149",
     INVALID_ID, dict(check_ascii=True, check_quality=True), [], "Line number"),
    ("ctx_mismatch", "tests/test_error_context.mojo:69-78",
     MISMATCH, dict(check_ascii=True, check_quality=True), [], "Record number"),
    ("ctx_record_number_2", "tests/test_error_context.mojo:97-137",
     SECOND_BAD, dict(check_ascii=True, check_quality=True), [(b"r1", b"AT", b"!@")],
     "Record number: 2"),
]

# batches(): (citation, data, batch_size, expected batch sizes)
BATCH_CASES = [
    ("tests/fastq/test_parser.mojo:122-138",
     b"@r1\nACGT\n+\n!!!!\n@r2\nTGCA\n+\n####\n@r3\nNNNN\n+\n!!!!\n", 2, [2, 1]),
    ("tests/fastq/test_parser.mojo:141-160",
     b"@a\nA\n+\n!\n@b\nB\n+\n!\n@c\nC\n+\n!\n@d\nD\n+\n!\n@e\nE\n+\n!\n", 2, [2, 2, 1]),
    ("tests/fastq/test_parser.mojo:180-192", b"", 4, []),
    ("tests/fastq/test_parser.mojo:195-209", b"@r1\nA\n+\n!\n", 4, [1]),
]

# tests/test_python_bindings.py:42,61-67
EXAMPLE_IDS = [b"EAS54_6_R1_2_1_413_324", b"EAS54_6_R1_2_1_540_792", b"EAS54_6_R1_2_1_443_348"]
