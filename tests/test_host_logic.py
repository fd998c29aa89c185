"""CPU-only tests of the host side: the C-ABI library loads and exports what the header
declares, configuration helpers, host containers and readers, shard stitching arithmetic, and the
N > 1 path over gloo (world_size 2).  No kernel runs here."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def B():
    import blazeseq_b200
    return blazeseq_b200


def test_library_exports_every_declared_symbol(B):
    hdr = open(os.path.join(ROOT, "include", "blazeseq_gpu.h")).read()
    declared = set(re.findall(r"\b(bsq_[a-z_0-9]+)\(", hdr))
    assert declared == set(B._capi.SYMBOLS), declared ^ set(B._capi.SYMBOLS)
    L = B._capi.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.bsq_abi_version() == 3


def test_struct_layouts_match_header(B):
    c = B._capi
    assert C.sizeof(c.Config) == 64 and C.sizeof(c.Error) == 1056 and C.sizeof(c.Summary) == 64
    assert C.sizeof(c.PassResult) == 48 + 1056 and C.sizeof(c.OffsetsView) == 48
    assert C.sizeof(c.BatchView) == 80 and C.sizeof(c.ShardStart) == 32


def test_no_cpu_fallback(B):
    """Without a CUDA device the product refuses to construct a parser."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B._capi.BsqLibraryError):
        B.FastqParser(B.MemoryReader(b"@a\nA\n+\n!\n"))
    h = C.c_void_p()
    assert B._capi.lib().bsq_create(None, C.byref(h)) == B._capi.E_NO_DEVICE


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "blazeseq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f in (), (f, "product code must not reference oracle/")


def test_config_defaults_and_schema_table(B, oracle):
    cfg = B._capi.default_config()
    assert (cfg.buffer_capacity, cfg.buffer_max_capacity, cfg.batch_size) == (256 * 1024, 1 << 30, 4096)
    assert (cfg.check_ascii, cfg.check_quality, cfg.buffer_growth_enabled) == (0, 0, 0)
    assert (cfg.q_lower, cfg.q_upper, cfg.q_offset) == (33, 126, 33)
    for name in ("generic", "sanger", "solexa", "illumina_1.3", "illumina_1.5", "illumina_1.8", "bogus"):
        assert B._capi.parse_schema(name) == oracle.schema(name)
    pc = B.ParserConfig()
    assert (pc.buffer_capacity, pc.check_ascii, pc.check_quality, pc.quality_schema) == (262144, False, False, None)


def test_synthetic_size_arithmetic(B, oracle):
    L = B._capi.lib()
    for t in (3 << 30, 10 << 30, 1 << 20, 0):
        for mn, mx in ((100, 100), (150, 150), (75, 300), (5, 12)):
            assert L.bsq_compute_num_reads_for_size(t, mn, mx) == oracle.compute_num_reads_for_size(t, mn, mx)
    for n, mn, mx in ((1000, 150, 150), (1000, 75, 300), (20, 5, 12), (1, 3, 3), (12345, 1, 500)):
        assert L.bsq_synth_size(n, mn, mx) == oracle.synth_size(n, mn, mx)
    assert L.bsq_synth_size(33659618, 150, 150) == 10737418142


def test_host_batch_container(B):
    """tests/fastq/test_record_batch.mojo:26-134 on the host-side FastqBatch."""
    batch = B.FastqBatch()
    batch.add(B.FastqRecord("a", "AC", "!!"))
    batch.add(B.FastqRecord("b", "GT", "!!"))
    assert batch.num_records() == 2 and batch.seq_len() == 4
    assert batch._ends.tolist() == [2, 4] and len(batch._quality_bytes) == 4 and len(batch._sequence_bytes) == 4
    recs = [B.FastqRecord("@a", "ACGT", "!!!!"), B.FastqRecord("@b", "TGCA", "!!!!"), B.FastqRecord("@c", "N", "!")]
    batch = B.FastqBatch()
    for r in recs:
        batch.add(r)
    assert batch.to_records() == recs and batch.get_record(2) == recs[2]
    assert batch.get_ref(1).sequence() == b"TGCA"
    with pytest.raises(B.BlazeSeqError):
        batch.get_record(3)
    assert recs[0].phred_scores == [0, 0, 0, 0] and len(recs[0]) == 4


def test_readers(B, tmp_path, golden_dir):
    data = b"@r1\nACGT\n+\n!!!!\n"
    r = B.MemoryReader(data)
    buf = np.zeros(10, np.uint8)
    assert r.read_to_buffer(buf, 10, 0) == 10 and bytes(buf) == data[:10]
    assert r.read_to_buffer(buf, 10, 0) == 6 and r.read_to_buffer(buf, 10, 0) == 0
    with pytest.raises(B.BlazeSeqError):
        r.read_to_buffer(buf, 11, 0)
    with pytest.raises(B.BlazeSeqError):
        r.read_to_buffer(buf, 1, 11)
    f = tmp_path / "x.fastq"
    f.write_bytes(data)
    fr = B.FileReader(f)
    big = np.zeros(64, np.uint8)
    assert fr.read_to_buffer(big, 64, 0) == len(data) and fr.read_to_buffer(big, 64, 0) == 0
    gz = B.RapidgzipReader(os.path.join(golden_dir, "corpus", "example.fastq.gz"))
    plain = open(os.path.join(golden_dir, "corpus", "example.fastq"), "rb").read()
    out = np.zeros(1024, np.uint8)
    n = gz.read_to_buffer(out, 1024, 0)
    assert bytes(out[:n]) == plain and gz.read_to_buffer(out, 1024, 0) == 0


# ------------------------------------------------------------------ shard stitching


def _tm():
    import oracle_py
    oracle_py.build()
    L = C.CDLL(os.path.join(ROOT, "oracle", "libbsq_tile_model.so"))
    L.tm_summarize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    return L


def _cpu_summary(B, tm, arr, lo, hi):
    s = B._capi.Summary()
    tm.tm_summarize(arr.ctypes.data, lo, hi, C.byref(s))
    return s


def test_shard_prefix_against_oracle(B, oracle):
    tm = _tm()
    rng = np.random.default_rng(0)
    streams = [oracle.synth(300, 5, 60, 2, 40, "sanger"), np.frombuffer(b"@a\nAC\n+\n!!\n" * 40, np.uint8),
               np.frombuffer(b"@a b\r\nAC\r\n+\r\n!!\r\n" * 30, np.uint8)]
    for data in streams:
        views, _, _ = oracle.parse_all(data)
        starts = views["header_start"]
        for _ in range(50):
            k = int(rng.integers(2, 9))
            cuts = np.sort(rng.integers(0, data.size + 1, k - 1))
            bounds = [0] + cuts.tolist() + [data.size]
            sums = [_cpu_summary(B, tm, data, bounds[i], bounds[i + 1]) for i in range(k)]
            st = B.shard_prefix(sums, [bounds[i + 1] - bounds[i] for i in range(k)])
            for i in range(k):
                first_own = int(np.searchsorted(starts, bounds[i]))
                assert st[i].first_record == first_own
                assert st[i].newline_rank == int((data[:bounds[i]] == 10).sum())
                owns = first_own < len(starts) and starts[first_own] < bounds[i + 1]
                if owns:
                    assert bounds[i] + st[i].skip_bytes == starts[first_own]
                else:
                    assert st[i].skip_bytes == bounds[i + 1] - bounds[i] or first_own == len(starts)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data_bytes, cuts, out_q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import blazeseq_b200 as B
    from blazeseq_b200 import sharding
    import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.frombuffer(data_bytes, np.uint8)
    bounds = [0] + list(cuts) + [data.size]
    lo, hi = bounds[rank], bounds[rank + 1]
    tm = _tm()
    # stand-in for bsq_summarize_device: the same BsqSummary, computed by the CPU tile model
    local = _cpu_summary(B, tm, data, lo, hi)
    plan = sharding.plan(dist, local, hi - lo)
    # stand-in for bsq_parse_device on the rank's own region (own shard + halo)
    region = data[lo + plan.begin: lo + plan.end]
    views, bases, err = O.parse_all(region, O.config(buffer_growth_enabled=True))
    reads, total_bases = sharding.allreduce_counts(dist, len(views), bases)
    out_q.put((rank, plan.first_record, len(views), reads, total_bases, err.code))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_parse_over_gloo(oracle, world):
    """world_size-2/3 on CPU: summaries all-gathered over gloo, each rank parses only the records
    that start in its shard, totals all-reduced.  The device steps are replaced by CPU stand-ins;
    the stitching (blazeseq_b200/sharding.py + bsq_shard_prefix) is the code under test."""
    import torch.multiprocessing as mp
    data = oracle.synth(5000, 20, 120, 2, 40, "illumina_1.8")
    views, bases, err = oracle.parse_all(data)
    rng = np.random.default_rng(world)
    cuts = sorted(int(x) for x in rng.integers(1, data.size - 1, world - 1))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data.tobytes(), cuts, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[2] for r in results) == len(views)
    first = 0
    for rank, first_record, n, reads, total_bases, code in results:
        assert first_record == first and code == oracle.EOF
        assert (reads, total_bases) == (len(views), bases)
        first += n


def test_bgzf_writer_members():
    """blazeseq_b200/bgzf.py writes what the stream pipeline's block-parallel inflater expects (SAM spec 4.1):
    gzip members of <= 64 KiB with the 'BC' extra field holding the member size, closed by the empty member;
    any gzip reader concatenates them back."""
    import gzip
    import struct

    from blazeseq_b200 import bgzf
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 200001, dtype=np.uint8).tobytes() + b"ACGT" * 40000
    for threads in (1, 4):
        blob = bgzf.compress(data, level=1, threads=threads)
        assert gzip.decompress(blob) == data
        pos, total, n = 0, 0, 0
        while pos < len(blob):
            assert blob[pos:pos + 4] == b"\x1f\x8b\x08\x04" and blob[pos + 10:pos + 16] == b"\x06\x00BC\x02\x00"
            size = struct.unpack_from("<H", blob, pos + 16)[0] + 1
            isize = struct.unpack_from("<I", blob, pos + size - 4)[0]
            assert isize <= 0x10000
            total += isize
            pos += size
            n += 1
        assert pos == len(blob) and total == len(data) and isize == 0 and n == -(-len(data) // bgzf.BLOCK) + 1


def test_host_batch_pipeline_sequencing_with_stand_in_parsers(B, oracle):
    """HostBatchPipeline's host logic (two handles alternating regions, the whole-batches cut, output offsets, error
    hand-over) with stand-in parser handles that answer from the oracle -- the device passes themselves are covered by
    tests/test_gpu_parity.py::test_whole_batches_regions_and_the_host_batch_pipeline."""
    import threading
    import time
    from types import SimpleNamespace
    from blazeseq_b200 import _capi as capi
    m = 250
    data = oracle.synth(9000, 40, 220, 2, 40, "sanger")
    views, bases, err = oracle.parse_all(data)
    exp = [oracle.build_batch(data, views[a:a + m]) for a in range(0, len(views), m)]
    active = {"n": 0, "max": 0}
    lock = threading.Lock()

    class StandIn:
        def __init__(self):
            self.cur = None
            self.closed = False

        def parse_host(self, region, stream_offset, first_record, is_last, want):
            assert want == capi.WANT_BATCHES | capi.WANT_WHOLE_BATCHES
            with lock:
                active["n"] += 1
                active["max"] = max(active["max"], active["n"])
            assert active["n"] == 1, "two regions were parsed at once: the cut of region k+1 needs region k's result"
            time.sleep(0.002)
            if not is_last:                       # a region that does not end the stream: its complete records only
                nl = np.flatnonzero(region == 10)
                k = len(nl) // 4 * 4
                region = region[:int(nl[k - 1]) + 1] if k else region[:0]
            v, _, e = oracle.parse_all(region, oracle.config(True, False, "sanger"))
            assert e.code == capi.EOF
            n = len(v)
            stop = SimpleNamespace(code=capi.EOF if is_last else capi.OK, text="", message=b"")
            if not is_last:
                n -= n % m
            consumed = region.size if is_last else (int(v[n]["header_start"]) if n < len(v) else region.size if n else 0)
            self.cur = (region, v[:n])
            with lock:
                active["n"] -= 1
            return SimpleNamespace(n_records=n, bytes_consumed=consumed, stop=stop)

        def soa_view(self):
            region, v = self.cur
            return SimpleNamespace(sequence_bytes=int(v["seq_len"].sum()), seq_len=int(v["qual_len"].sum()),
                                   total_id_bytes=int(v["id_len"].sum()))

        def soa_to_host(self, seq, qual, idb, ends, id_ends):
            region, v = self.cur
            time.sleep(0.003)                     # the other handle parses the next region meanwhile
            so = qo = io = ro = 0
            for a in range(0, len(v), m):
                i_, s_, q_, ie_, e_ = oracle.build_batch(region, v[a:a + m])
                seq[so:so + s_.size] = s_; qual[qo:qo + q_.size] = q_; idb[io:io + i_.size] = i_
                ends[ro:ro + e_.size] = e_; id_ends[ro:ro + ie_.size] = ie_
                so += s_.size; qo += q_.size; io += i_.size; ro += e_.size

        def close(self):
            self.closed = True

    tot = {k: sum(e[i].size for e in exp) for k, i in (("id", 0), ("seq", 1), ("qual", 2))}
    for region_bytes in (90_000, 400_000, 10 ** 9):
        pipe = B.HostBatchPipeline(StandIn, region_bytes=region_bytes)
        seq = np.zeros(tot["seq"], np.uint8); qual = np.zeros(tot["qual"], np.uint8); idb = np.zeros(tot["id"], np.uint8)
        ends = np.zeros(len(views), np.int64); id_ends = np.zeros(len(views), np.int64)
        got = pipe.run(data, seq, qual, idb, ends, id_ends)
        assert got == (len(views), tot["seq"], tot["qual"], tot["id"]) and pipe.stop.code == capi.EOF
        assert np.array_equal(seq, np.concatenate([e[1] for e in exp])) and np.array_equal(qual, np.concatenate([e[2] for e in exp]))
        assert np.array_equal(idb, np.concatenate([e[0] for e in exp]))
        assert np.array_equal(ends, np.concatenate([e[4] for e in exp])) and np.array_equal(id_ends, np.concatenate([e[3] for e in exp]))
        parsers = list(pipe.parsers)
        pipe.close()
        assert all(p.closed for p in parsers)
    # a region smaller than one batch cannot make progress: reported, not looped on
    pipe = B.HostBatchPipeline(StandIn, region_bytes=5_000)
    with pytest.raises(ValueError):
        pipe.run(data, None, None, None, None, None)
    pipe.close()


def _build_cpp_runner(tmp_path):
    import subprocess
    exe = os.path.join(tmp_path, "run_blazeseq")
    lib = os.path.join(ROOT, "blazeseq_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "run_blazeseq.cpp"), "-L" + lib, "-lblazeseq_gpu", "-Wl,-rpath," + lib, "-o", exe],
                   check=True, capture_output=True)
    return exe


def test_cpp_host_mirror_builds_and_never_parses_on_the_cpu(B, tmp_path, golden_dir):
    """include/blazeseq_gpu.hpp (C++ host mirror over the C ABI) + examples/run_blazeseq.cpp (the reference's benchmark
    runners) compile warning-free against the library; without a device the runner reports that and parses nothing."""
    import subprocess
    B._capi.lib()      # the library is built
    exe = _build_cpp_runner(str(tmp_path))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([exe, os.path.join(golden_dir, "corpus", "example.fastq")], capture_output=True, text=True)
    if has_gpu:
        assert r.returncode == 0 and r.stdout.split()[0] == "3", (r.stdout, r.stderr)
    else:
        assert r.returncode != 0 and r.stdout.split() == ["0", "0"] and "no usable CUDA device" in r.stderr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the keys of the
    bench contract; it needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gib", "0.02", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert r.returncode == 0 and len(lines) == 1, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "fastq_reads_per_s" and d["unit"] == "reads/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
