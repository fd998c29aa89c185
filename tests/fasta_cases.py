"""Literal FASTA streams and expectations transcribed from the reference's OWN tests
(tests/fasta/test_fasta_parser.mojo; line numbers cited per case).  Each case: (citation, data, check_ascii,
expected records [(id, sequence)], substring of the terminal error -- "EOF" for a clean end)."""

CASES = [
    ("test_fasta_parser.mojo:92-110 single line", b">id1\nACGT\n", False, [(b"id1", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:113-125 multi-line", b">id1\nACG\nTTA\nGG\n", False, [(b"id1", b"ACGTTAGG")], "EOF"),
    ("test_fasta_parser.mojo:128-144 back to back", b">id1\nACGT\n>id2\nTTAA\n", False,
     [(b"id1", b"ACGT"), (b"id2", b"TTAA")], "EOF"),
    ("test_fasta_parser.mojo:147-158 no terminal newline", b">id1\nACGT", False, [(b"id1", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:161-177 first line not a header", b"ACGT\n>id1\nACGT\n", False, [], "does not start with"),
    ("test_fasta_parser.mojo:208-217 non-ASCII id", b">id\x80\nACGT\n", True, [], "Non ASCII"),
    ("test_fasta_parser.mojo:219-228 non-ASCII sequence", b">id1\nAC\x80T\n", True, [], "Non ASCII"),
    ("test_fasta_parser.mojo:230-246 iterator", b">id1\nAC\nGT\n>id2\nTT\nAA\n", False,
     [(b"id1", b"ACGT"), (b"id2", b"TTAA")], "EOF"),
    ("test_fasta_parser.mojo:336-364 five records", b">alpha\nAAAA\n>beta\nCCCC\n>gamma\nGGGG\n>delta\nTTTT\n>epsilon\nACGT\n",
     False, [(b"alpha", b"AAAA"), (b"beta", b"CCCC"), (b"gamma", b"GGGG"), (b"delta", b"TTTT"), (b"epsilon", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:433-452 six lines", b">id1\nACGTACGTAC\nGTACGTACGT\nACGTACGTAC\nGTACGTACGT\nACGTACGTAC\nGTACGTACGT\n",
     False, [(b"id1", b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT")], "EOF"),
    ("test_fasta_parser.mojo:527-535 leading blank lines", b"\n\n\n>id1\nACGT\n", False, [(b"id1", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:538-550 blank lines between records", b">id1\nACGT\n\n\n>id2\nTTAA\n", False,
     [(b"id1", b"ACGT"), (b"id2", b"TTAA")], "EOF"),
    ("test_fasta_parser.mojo:553-562 CRLF", b">id1\r\nACGT\r\n", False, [(b"id1", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:565-575 CRLF two records", b">id1\r\nACGT\r\n>id2\r\nTTAA\r\n", False,
     [(b"id1", b"ACGT"), (b"id2", b"TTAA")], "EOF"),
    ("test_fasta_parser.mojo:578-589 id leading blanks", b">  spaced_id\nACGT\n", False, [(b"spaced_id", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:592-603 id trailing blanks", b">seq_id   \nACGT\n", False, [(b"seq_id", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:606-617 id tabs", b">\ttab_id\t\nACGT\n", False, [(b"tab_id", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:620-628 empty id", b">\nACGT\n", False, [(b"", b"ACGT")], "EOF"),
    ("test_fasta_parser.mojo:631-639 one base", b">id1\nA\n", False, [(b"id1", b"A")], "EOF"),
    ("test_fasta_parser.mojo:642-659 lower / mixed case", b">id1\nAcGtAcGt\n", False, [(b"id1", b"AcGtAcGt")], "EOF"),
    ("test_fasta_parser.mojo:662-673 single-base lines", b">id1\nA\nC\nG\nT\nA\nC\nG\nT\n", False, [(b"id1", b"ACGTACGT")], "EOF"),
    ("test_fasta_parser.mojo:676-690 multi-line, no terminal newline", b">id1\nACG\nTTA", False, [(b"id1", b"ACGTTA")], "EOF"),
    ("test_fasta_parser.mojo:714-731 empty sequence at EOF", b">id1\n", False, [], "empty sequence"),
    ("test_fasta_parser.mojo:734-754 empty sequence before a header", b">id1\n>id2\nACGT\n", False, [], "empty sequence"),
    ("test_fasta_parser.mojo:757-771 empty file", b"", False, [], "EOF"),
    ("test_fasta_parser.mojo:774-788 blanks only", b"\n\n   \n\t\n", False, [], "EOF"),
    ("test_fasta_parser.mojo:791-807 no header at all", b"ACGTACGT\n", False, [], "does not start with"),
    ("test_fasta_parser.mojo:810-827 second record empty", b">id1\nACGT\n>id2\n>id3\nGGGG\n", False, [(b"id1", b"ACGT")],
     "empty sequence"),
]

# tests/fasta/test_fasta_parser_correctness.mojo (Biopython files under tests/test_data/fasta_parser/):
# (file, record count or None for ">= 1", [(index, id or id substring, sequence or substring or None)])
CORPUS = [
    ("f002", 3, [(0, b"gi|1348912|gb|G26680|", b"CGGACCAGACGGACACAGGGAGAAGCTAGTTTCTTTCATGTGATTGA"), (2, b"gi|1592936|gb|G29385|", None)]),
    ("f003.fa", 2, [(0, b"gi|3318709|pdb|1A91|", b"MENLNMDLLYMAAAVMMGLAAIGAAIGIGILGGKFLEGAARQPDLIPLLRTQFFIVMGLVDAIPMIAVGLGLYVMFAVA"),
                    (1, b"gi|whatever|whatever", b"MENLNMDLLYMAAAVMMGLAAIGAAIGIGILGG")]),
    ("fa01", 2, [(0, b"AK1H_ECOLI/1-378", b"CPDSINAALICRGEKMSIAIMAGVLEARGH"), (1, b"AKH_HAEIN/1-382", b"VEDAVKATIDCRGEKLSIAMMKAWFEARGY")]),
    ("aster.pro", None, [(0, b"gi|3298468|dbj|BAA31520.1|", None)]),
    ("aster_no_wrap.pro", None, [(0, b"gi|3298468|dbj|BAA31520.1|", None)]),
]
