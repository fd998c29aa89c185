"""Small end-to-end inputs for compute-sanitizer (scripts/gpu_sanitize.sh): the device inflater on the reference's .bgz
fixtures and a synthetic BGZF file (all block types), the FASTA path, a 320-byte-stride batch pass (rotated staging), the
device writer."""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import blazeseq_b200 as B  # noqa: E402
from blazeseq_b200 import _capi as capi, bgzf  # noqa: E402
import oracle_py as O  # noqa: E402

tmp = tempfile.mkdtemp()
gold = os.path.join(ROOT, "tests", "golden")
rng = np.random.default_rng(1)
data = O.synth(10 ** 8 + 1, 150, 150, 2, 40, "sanger", first=5, count=1500).tobytes()        # 320-byte records
blobs = {"fixture": open(os.path.join(gold, "corpus", "example.fastq.bgz"), "rb").read(),
         "level6": bgzf.compress(data, 6), "level0": bgzf.compress(data[:200000], 0),
         "noise": bgzf.compress(rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(), 6)}
for name, blob in blobs.items():
    path = os.path.join(tmp, name + ".bgz")
    open(path, "wb").write(blob)
    g = B.GpuParser(batch_size=256)
    st = g.stream_open(path, capi.SOURCE_GZIP, 64 << 20)
    res, region, off, first = g.stream_next(st, capi.WANT_OFFSETS | capi.WANT_BATCHES)
    exp, rest = b"", blob
    while rest:
        d = zlib.decompressobj(31)
        exp += d.decompress(rest)
        rest = d.unused_data
    assert bytes(region) == exp, name
    g.stream_close(st)
    g.close()
g = B.GpuParser(check_ascii=True, batch_size=512)
arr = np.frombuffer(data, np.uint8)
res = g.parse_host(arr, want=3)
views, bases, err = O.parse_all(arr, O.config(True, False))
assert res.n_records == len(views) == 1500 and res.n_bases == bases
# the device writer (FastqRecord.write over the SoA) and the whole-batches cut
res = g.parse_host(arr, want=capi.WANT_BATCHES)
text, offs = g.write_records(want_offsets=True)
assert text.tobytes() == data and int(offs[-1]) == len(data)
res = g.parse_host(arr[:300000], 0, 0, False, capi.WANT_BATCHES | capi.WANT_WHOLE_BATCHES)
assert res.n_records % 512 == 0 and res.n_records > 0
fa = b">a desc\nACGT\nAC GT\r\n\n>b\nTTTT" * 50
r = g.fasta_parse_host(np.frombuffer(fa, np.uint8))
ids, seqs, e = O.fasta_parse(fa, True)
assert r.n_records == len(ids) and r.stop.code == e.code
g.close()
print("sanitize_small ok")
