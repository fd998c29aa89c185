#!/usr/bin/env python
"""BASELINE configs[4] (scaled): a .fastq.gz streamed through the native pipeline (bsq_stream_*):
reader thread (zlib inflate, the GZFile / RapidgzipReader(parallelism=0) role) -> pinned regions ->
H2D on the copy stream -> scan/resolve/pack.  Prints one JSON line with the overlap accounting.

    python scripts/bench_gzip.py [--gib 1.0] [--region-mib 256]
"""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import blazeseq_b200 as B
from blazeseq_b200 import _capi as capi

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=1.0, help="uncompressed payload")
ap.add_argument("--region-mib", type=int, default=256)
args = ap.parse_args()

schema = B.parse_schema("illumina_1.8")
gpu = B.GpuParser(False, False, schema, 4096)
M = capi.lib().bsq_compute_num_reads_for_size(int(args.gib * (1 << 30)), 150, 150)
size = capi.lib().bsq_synth_size(M, 150, 150)
buf = torch.empty(size + 256, dtype=torch.uint8, device="cuda:0")
gpu.synth_device(buf.data_ptr(), size, M, 0, M, 150, 150, 2, 40, schema)
host = buf[:size].cpu().numpy()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
plain, gz = os.path.join(tmp, "x.fastq"), os.path.join(tmp, "x.fastq.gz")
host.tofile(plain)
# multi-member gzip written by all cores (members are cut at arbitrary bytes; gzread concatenates them)
step = 32 << 20
chunks = [host[i:i + step] for i in range(0, size, step)]
with ThreadPoolExecutor(os.cpu_count()) as ex:
    parts = list(ex.map(lambda c: gzip.compress(c.tobytes(), compresslevel=6), chunks))
with open(gz, "wb") as f:
    for p in parts:
        f.write(p)
gz_size = os.path.getsize(gz)
# the same payload as BGZF (64 KiB members): inflated block-parallel by the reader's worker threads
from blazeseq_b200 import bgzf
bgz = os.path.join(tmp, "x.fastq.bgz")
with open(bgz, "wb") as f:
    f.write(bgzf.compress(host, 6, threads=os.cpu_count()))
bgz_size = os.path.getsize(bgz)


def run(path):
    st = gpu.stream_open(path, capi.SOURCE_AUTO, args.region_mib << 20)
    t0 = time.perf_counter()
    recs = 0
    while True:
        res, region, off, first = gpu.stream_next(st, capi.WANT_BATCHES)
        recs += int(res.n_records)
        if res.stop.code != capi.OK:
            assert res.stop.code == capi.EOF, res.stop.text
            break
    wall = time.perf_counter() - t0
    s = gpu.stream_stats(st)
    gpu.stream_close(st)
    assert recs == M
    return {"wall_s": wall, "reads_per_s": recs / wall, "uncompressed_gb_per_s": size / wall / 1e9,
            "reader_busy_s": s.reader_busy_s, "gpu_pass_s": s.parse_s, "caller_wait_reader_s": s.wait_reader_s,
            "regions": int(s.regions), "overlap": "wall ~= max(reader, passes): %.2f vs %.2f + %.2f" %
            (wall, s.reader_busy_s, s.parse_s)}


run(plain)  # warm the page cache and the arenas
out = {"workload": "configs[4] scaled: %.2f GiB uncompressed 150 bp FASTQ (%d reads), gzip -6, %.2f GiB compressed; "
                   "single-thread zlib inflate in the reader thread (RapidgzipReader(parallelism=0) role), batches(4096)"
                   % (size / (1 << 30), M, gz_size / (1 << 30)),
       "region_mib": args.region_mib, "gzip": run(gz), "bgzf_all_threads": run(bgz), "plain_file": run(plain),
       "bgzf_compressed_gib": bgz_size / (1 << 30), "host_threads": os.cpu_count()}
print(json.dumps(out))
os.remove(plain); os.remove(gz); os.remove(bgz); os.rmdir(tmp)
