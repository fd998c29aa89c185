#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

What it writes
  corpus/*                 the reference's FASTQ parser fixtures (data files, not source):
                           tests/test_data/fastq_parser/*.fastq|*.gz|*.bgz
  corpus_expect.json       per file, taken from the reference's OWN tests and docs:
                             readme_current_error  <- tests/test_data/fastq_parser/README.md:7-75
                             invalid_msg           <- tests/fastq/test_fastq_parser_correctness.mojo:511-747
                             valid_schema          <- same file, valid_file_test_fun(...) calls :141-509
                           plus the accept-set the reference's invalid tests use (:21-56).
  synthetic_kat.json       SHA-256 / size known answers of the synthetic generator
                           (utils.mojo:831-917) produced by the oracle restatement and equal to
                           the values SURVEY.md App. B.3 derived independently.

/root/reference does not exist on the GPU box, so everything the tests need is committed here.
"""
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(REF, "tests/test_data/fastq_parser")
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))


def main():
    os.makedirs(os.path.join(HERE, "corpus"), exist_ok=True)
    files = sorted(f for f in os.listdir(SRC) if f.endswith((".fastq", ".gz", ".bgz")))
    for f in files:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(HERE, "corpus", f))

    readme = {}
    for line in open(os.path.join(SRC, "README.md"), encoding="utf-8"):
        m = re.match(r"\|\s*([\w\-.]+\.fastq)\s*\|(.*)\|(.*)\|(.*)\|\s*$", line)
        if m:
            readme[m.group(1)] = m.group(4).strip()

    test_src = open(os.path.join(REF, "tests/fastq/test_fastq_parser_correctness.mojo"),
                    encoding="utf-8").read()
    consts = dict(re.findall(r'comptime (\w+) = "([^"]*)"', test_src))
    consts["EOF"] = "EOF"
    invalid = {}
    for f, msg in re.findall(r'invalid_file_test_fun\("([^"]+)",\s*(\w+)\)', test_src):
        invalid[f] = consts[msg]
    valid = {}
    for f, schema in re.findall(r'\bvalid_file_test_fun(?:_ref)?\(\s*"([^"]+)"(?:,\s*"([^"]+)")?\s*\)',
                                test_src):
        valid[f] = schema or "generic"
    valid_gz = {}
    for f, schema in re.findall(r'valid_file_test_fun_gz(?:_ref)?\(\s*"([^"]+)"(?:,\s*"([^"]+)")?\s*\)',
                                test_src):
        valid_gz[f] = schema or "generic"
    # multi-line files: tests are commented out in the reference (README "Multi-line (disabled)")
    disabled = ["tricky.fastq", "longreads_original_sanger.fastq", "wrapping_original_sanger.fastq"]

    expect = {
        "source": "MoSafi2/BlazeSeq @ 66ddbd1: tests/test_data/fastq_parser/README.md, "
                  "tests/fastq/test_fastq_parser_correctness.mojo",
        "invalid_config": {"check_ascii": True, "check_quality": True, "schema": "generic"},
        "accept_set": [consts["cor_len"], consts["cor_seq_hed"], consts["plus_line_start"],
                       consts["sep_line_start"], "EOF"],
        "multi_line_disabled": disabled,
        "files": {},
    }
    for f in files:
        if not f.endswith(".fastq"):
            continue
        expect["files"][f] = {
            "readme_current_error": readme.get(f),
            "invalid_msg": invalid.get(f),
            "valid_schema": valid.get(f),
        }
    expect["gz_files"] = valid_gz
    json.dump(expect, open(os.path.join(HERE, "corpus_expect.json"), "w"), indent=1, sort_keys=True)

    import oracle_py as O
    kat = {"source": "oracle restatement of utils.mojo:831-917; equal to SURVEY.md App. B.3", "cases": []}
    for args in [(1000, 150, 150, 2, 40, "sanger"), (1000, 75, 300, 2, 40, "illumina_1.8"),
                 (20, 5, 12, 2, 25, "generic"), (12, 5, 11, 2, 40, "generic"),
                 (1000, 50, 150, 2, 40, "generic")]:
        b = O.synth(*args)
        kat["cases"].append({"args": list(args), "bytes": int(b.size), "sha256": O.sha256(b),
                             "head": bytes(b[:64]).decode("latin-1")})
    kat["compute_num_reads_for_size"] = [
        {"args": [t, mn, mx], "reads": O.compute_num_reads_for_size(t, mn, mx)}
        for t in (3 << 30, 10 << 30) for mn, mx in ((100, 100), (150, 150), (75, 300))]
    json.dump(kat, open(os.path.join(HERE, "synthetic_kat.json"), "w"), indent=1)
    print("wrote", len(files), "corpus files,", len(expect["files"]), "expectations")


if __name__ == "__main__":
    main()
