#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200.

Metric (BASELINE.json): FASTQ reads/s (and parsed GB/s) of the scan / validate / FastqBatch-pack hot
path.  Workload at every N: BASELINE.json configs[1] -- 10 GiB in-memory 150 bp Illumina FASTQ
(33,659,618 reads, generate_synthetic_fastq_buffer semantics), schema illumina_1.8, validation OFF,
batches(4096) -- PER GPU (weak scaling: rank r holds records [r*M, (r+1)*M) of an N*M-record stream).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path
    python bench.py --mixed --shard-stream                    # configs[3]: ONE 10 GiB mixed-length stream cut
                                                              # at arbitrary byte offsets over the N ranks

One "step" = one pass of the hot path over the whole per-GPU input, already resident in HBM
(`value`), or starting from pinned HOST memory through bsq_parse_host (`e2e`).  The default run also
reports configs[2] (validation on), views() and configs[3] (sharded mixed stream) as `sub_results`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class Stream:
    """Geometry of generate_synthetic_fastq_buffer(reads, mn, mx, ...) (utils.mojo:753-768): len_i = mn + (31 i + 7) %
    (mx - mn + 1), header "@read_<i zero padded to the width of reads - 1>\\n"."""

    def __init__(self, reads: int, mn: int, mx: int):
        self.reads, self.mn, self.mx = reads, mn, mx
        self.digits = len(str(reads - 1)) if reads > 1 else 1
        self.period = mx - mn + 1
        self.pref = [0]
        for j in range(self.period):
            self.pref.append(self.pref[-1] + (31 * j + 7) % self.period)
        self.overhead = 6 + self.digits + 1 + 4          # '@read_' + digits + 4 line ends + '+'
        self.id_len = 5 + self.digits

    def offset(self, i: int) -> int:
        lens = i * self.mn + (i // self.period) * self.pref[self.period] + self.pref[i % self.period]
        return i * self.overhead + 2 * lens

    def bases(self, a: int, b: int) -> int:
        return (self.offset(b) - self.offset(a) - (b - a) * self.overhead) // 2

    def record_at(self, byte: int) -> int:
        """Largest i with offset(i) <= byte."""
        lo, hi = 0, self.reads
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if self.offset(mid) <= byte:
                lo = mid
            else:
                hi = mid - 1
        return lo


def workload_config(args, world: int, M: int, size: int, rec_bytes: float) -> dict:
    """The `config` object of the JSON line: identical for this repo's arm and the reference arm."""
    if args.shard_stream:
        name = ("configs[3]: ONE %g GiB mixed read-length (75-300 bp) FASTQ stream, validation OFF" if args.mixed
                else "ONE %g GiB 150 bp FASTQ stream") % args.gib
        return {"workload": name + f", {args.mode}(4096), cut at arbitrary byte offsets over the ranks",
                "reads_total": M, "bytes_total": size, "record_bytes": rec_bytes,
                "l2": "input (>=10 GB) is larger than L2; no flush needed",
                "parallelism": f"{world} contiguous byte shards; per step: shard summary (k_summarize) -> all-gather of 72 B "
                               f"-> bsq_shard_prefix -> own records + halo -> all-reduce of counts"}
    name = ("configs[3]: 10 GiB mixed read-length (75-300 bp) FASTQ, validation OFF" if args.mixed
            else "configs[2]: 10 GiB 150 bp FASTQ, check_ascii+check_quality, sanger" if args.validate
            else "configs[1]: 10 GiB in-memory 150 bp Illumina FASTQ, illumina_1.8, validation OFF")
    if args.crlf:
        name += ", CRLF line ends"
    return {"workload": name + f", {args.mode}(4096), per GPU", "reads_per_gpu": M, "bytes_per_gpu": size,
            "record_bytes": rec_bytes, "l2": "input (>=10 GB) is larger than L2; no flush needed",
            "parallelism": f"{world} x record-aligned shard, NCCL all-reduce of counts only"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def reference_arm(args):
    """CPU restatement of the reference path (the oracle port: Mojo cannot be built in this image),
    all host threads, on this arm's workload: the rank-0 shard of the same stream, whole.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py as O
    from concurrent.futures import ThreadPoolExecutor

    O.build()
    world = max(1, args.gpus)
    cores = os.cpu_count() or 1
    mn, mx = (75, 300) if args.mixed else (args.read_len, args.read_len)
    M = O.compute_num_reads_for_size(int(args.gib * GIB), mn, mx)
    st = Stream(M if args.shard_stream else M * world, mn, mx)
    first = 0                                       # rank 0's shard: records [0, M)
    full_size = st.offset(first + M) - st.offset(first)
    sample_reads = M
    if args.ref_sample_gib > 0:
        sample_reads = max(1, min(M, st.record_at(int(args.ref_sample_gib * GIB))))
    size = st.offset(first + sample_reads) - st.offset(first)
    data = np.empty(size, np.uint8)
    parts = max(1, min(cores, 64))
    step = (sample_reads + parts - 1) // parts

    def gen(i):
        a = first + i * step
        cnt = min(step, first + sample_reads - a)
        if cnt > 0:
            lo, hi = st.offset(a) - st.offset(first), st.offset(a + cnt) - st.offset(first)
            if mn == mx:
                O.synth(st.reads, mn, mx, 2, 40, "illumina_1.8", first=a, count=cnt, out=data[lo:hi])
            else:   # (the generator wants room for `cnt` records of the maximum length)
                data[lo:hi] = O.synth(st.reads, mn, mx, 2, 40, "illumina_1.8", first=a, count=cnt)
    with ThreadPoolExecutor(parts) as ex:
        list(ex.map(gen, range(parts)))
    cfg = O.config(args.validate, args.validate, "sanger" if args.validate else "illumina_1.8")
    mode = 1 if args.mode == "batches" else 0
    for _ in range(args.warmup):
        O.baseline_mt(data, cfg, mode, 4096, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, bases, code = O.baseline_mt(data, cfg, mode, 4096, cores)
        assert n == sample_reads and code == 0
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_reads / dt
    # the reference's FastqParser is a sequential object (one parser = one thread, record.mojo:439): the same
    # bytes through ONE thread is what a single BlazeSeq parser delivers; `value` shards them by newline rank
    # over every host thread, which the reference itself does not do
    t1 = time.perf_counter()
    O.baseline_mt(data, cfg, mode, 4096, 1)
    one_thread = sample_reads / (time.perf_counter() - t1)
    whole = sample_reads == M
    sample = (f"{'all' if whole else 'the first'} {sample_reads} reads ({data.size / GIB:.2f} GiB) of the workload"
              f"{'' if whole else ' (bounded sample)'}, {args.mode}(4096), {cores} threads")
    emit(json.dumps({
        "impl": "reference", "metric": "fastq_reads_per_s", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.shard_stream else "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "parsed_gb_per_s": data.size / dt / 1e9,
        "config": workload_config(args, world, M, full_size, full_size / M),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample,
                         "value_1core": one_thread},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pin_to_gpu_numa_node(index: int) -> dict:
    """Binds this process to the CPUs of the NUMA node the GPU hangs off (PCI sysfs), so that its pinned buffers and
    reader threads are local to the GPU's root port.  Returns what it found (reported in the e2e object)."""
    info = {"numa_node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                      # sysfs uses a 4-digit PCI domain
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read())
        cpulist = open(f"{base}/local_cpulist").read().strip()
        info["numa_node"] = node
        info["cpus"] = cpulist
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        if cpus and node >= 0:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            info["bound"] = True
    except Exception as e:   # (no sysfs entry, a container without the topology): nothing to bind to
        info["note"] = repr(e)[:120]
    return info


class _DevPtr:
    """A device address as a torch-importable array (zero copy)."""

    def __init__(self, ptr, n, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="blazeseq_b200")
    ap.add_argument("--gib", type=float, default=10.0, help="per-GPU input size target (default: the 10 GiB config)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the configs[2] / views / configs[3] sub-results")
    ap.add_argument("--sub-steps", type=int, default=3)
    ap.add_argument("--ref-sample-gib", type=float, default=0.0, help="reference arm: bound the input (0 = the whole workload)")
    ap.add_argument("--cpu-sample-gib", type=float, default=1.0)
    ap.add_argument("--mode", default="batches", choices=["batches", "views"])
    ap.add_argument("--validate", action="store_true", help="configs[2]: check_ascii + check_quality, sanger")
    ap.add_argument("--mixed", action="store_true", help="configs[3]: mixed read length 75-300 bp instead of 150 bp")
    ap.add_argument("--shard-stream", action="store_true",
                    help="strong scaling: ONE stream of --gib cut at arbitrary byte offsets over the ranks (sharding.plan)")
    ap.add_argument("--gzip", action="store_true", help="configs[4]: .fastq.gz / BGZF files through the stream pipeline (--gib 4)")
    ap.add_argument("--region-mib", type=int, default=512, help="--gzip: region size of the stream pipeline")
    ap.add_argument("--crlf", action="store_true", help="CRLF line ends: every id needs _strip_spaces (id strip pipeline)")
    ap.add_argument("--read-len", type=int, default=150, help="read length of the synthetic stream (record stride sweeps)")
    ap.add_argument("--id-digits", type=int, default=0, help="zero-padded id width (0: that of the stream's read count)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch

    import blazeseq_b200 as B
    from blazeseq_b200 import _capi as capi
    from blazeseq_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    host_info = pin_to_gpu_numa_node(local)   # before any pinned allocation: first touch decides where it lives
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    L = capi.lib()
    peak, peak_src = measured_peaks()

    def timed(step_fn, steps, gpu):
        """`steps` calls between barrier+synchronize pairs.  Returns (wall s, max over ranks; per-step device ms [5];
        per-step device ms of the whole pass, max over ranks; kernel launches)."""
        barrier()
        ms_sum, launches = [0.0] * 5, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
            ms, nl = gpu.timing()
            ms_sum = [a + b for a, b in zip(ms_sum, ms)]
            launches += nl
        barrier()
        wall = time.perf_counter() - t0
        wall_max, dev_ms_max, launches = max_over_ranks(wall, ms_sum[4] / steps, float(launches))
        return wall_max, [m / steps for m in ms_sum], dev_ms_max, int(launches)

    # ------------------------------------------------------------------------------------------------
    # configs[3] / --shard-stream: ONE stream, contiguous byte shards cut at arbitrary offsets
    # ------------------------------------------------------------------------------------------------
    def shard_stream_leg(gib, mn, mx, steps, warmup, mode="batches"):
        schema = B.parse_schema("illumina_1.8")
        reads = L.bsq_compute_num_reads_for_size(int(gib * GIB), mn, mx)
        st = Stream(reads, mn, mx)
        total = st.offset(reads)
        # cut points: equal shares, moved off every natural alignment
        bounds = [0] + [total * r // world + 37 * r + 5 for r in range(1, world)] + [total]
        lo, hi = bounds[rank], bounds[rank + 1]
        halo = 2 * (st.overhead + 2 * mx)                       # the next shard's first own record starts within one record
        end = min(total, hi + halo)
        i0, i1 = st.record_at(lo), min(reads, st.record_at(end - 1) + 1)
        nbytes = st.offset(i1) - st.offset(i0)
        tmp = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        g = B.GpuParser(False, False, schema, 4096, device_id=local)
        assert g.synth_device(tmp.data_ptr(), nbytes, reads, i0, i1 - i0, mn, mx, 2, 40, schema) == nbytes
        base = tmp.data_ptr() + (lo - st.offset(i0))             # device address of stream byte `lo`
        want = capi.WANT_BATCHES if mode == "batches" else capi.WANT_OFFSETS
        first_own = i0 if st.offset(i0) == lo else i0 + 1        # what the stitching must find
        state = {}

        def step():
            summary = g.summarize_device(base, hi - lo)
            if dist is not None:
                plan = sharding.plan(dist, summary, hi - lo, device=dev)
            else:
                plan = sharding.ShardPlan(0, 1, 0, hi - lo, 0, 0)
            assert plan.first_record == first_own and plan.end - (hi - lo) <= end - hi, (plan, first_own)
            n = plan.end - plan.begin
            r = g.parse_device(base + plan.begin, n, lo + plan.begin, plan.first_record, rank == world - 1, want)
            assert r.bytes_consumed == n and r.stop.code in (capi.OK, capi.EOF), (r.bytes_consumed, n, r.stop.text)
            state["r"], state["plan"] = r, plan
            return r
        for _ in range(max(warmup, 1)):
            step()
        wall, ms, dev_ms, launches = timed(step, steps, g)
        r, plan = state["r"], state["plan"]
        own_reads = int(r.n_records)
        own_bases = st.bases(plan.first_record, plan.first_record + own_reads)
        assert r.n_bases == own_bases
        tot_reads, tot_bases = own_reads, own_bases
        if dist is not None:
            tot_reads, tot_bases = sharding.allreduce_counts(dist, own_reads, own_bases, device=dev)
        assert (tot_reads, tot_bases) == (reads, st.bases(0, reads)), (tot_reads, reads)
        g.close()
        del tmp
        algo = total / reads + (2 * st.bases(0, reads) / reads + st.id_len + 16 if mode == "batches" else 20)
        return {"reads": reads, "bytes": total, "record_bytes": total / reads, "wall": wall, "ms": ms, "dev_ms": dev_ms,
                "launches": launches, "steps": steps, "value": reads / (wall / steps), "ms_per_step": wall / steps * 1e3,
                "n_windows": int(r.n_windows), "own_reads": own_reads, "algo": algo}

    # ------------------------------------------------------------------------------------------------
    # configs[4]: a .fastq.gz through the native stream pipeline (bsq_stream_*), file -> results
    # ------------------------------------------------------------------------------------------------
    def gzip_leg(gib, region_mib=512):
        """Writes `gib` of the 150 bp stream as a plain file, as BGZF (gzip level 6, 64 KiB members) and as an ordinary
        single-member gzip file, and streams each through bsq_stream_next(WANT_BATCHES).  BGZF: the compressed members
        cross PCIe and are inflated on the device (k_inflate_members); ordinary gzip is decoded speculatively in parallel
        by the host threads (bsq_pgzip.h), and by one zlib thread for comparison.  The CPU arm beside it: zlib on every
        host thread (BGZF members) feeding nothing (inflate only)."""
        import gzip as gz
        import shutil
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        from blazeseq_b200 import bgzf
        schema = B.parse_schema("illumina_1.8")
        g = B.GpuParser(False, False, schema, 4096, device_id=local)
        gh = B.GpuParser(False, False, schema, 4096, device_id=local, host_inflate=True)
        reads = L.bsq_compute_num_reads_for_size(int(gib * GIB), 150, 150)
        nbytes = L.bsq_synth_size(reads, 150, 150)
        dbuf = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        g.synth_device(dbuf.data_ptr(), nbytes, reads, 0, reads, 150, 150, 2, 40, schema)
        host = dbuf[:nbytes].cpu().numpy()
        del dbuf
        tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        cores = os.cpu_count() or 1
        try:
            plain, gzp, bgz = (os.path.join(tmp, n) for n in ("x.fastq", "x.fastq.gz", "x.fastq.bgz"))
            host.tofile(plain)
            # ordinary gzip: ONE member (what `gzip` / `pigz` write; no member boundaries to split at), deflated by all
            # cores the way pigz does it: 32 MiB pieces of raw deflate ending in a sync flush, the last one finished
            step = 32 << 20
            starts = list(range(0, nbytes, step))

            def deflate_piece(i):
                co = zlib.compressobj(6, zlib.DEFLATED, -15)
                z = co.compress(host[i:i + step].tobytes())
                return z + co.flush(zlib.Z_FINISH if i == starts[-1] else zlib.Z_SYNC_FLUSH)
            with ThreadPoolExecutor(cores) as ex:
                parts = list(ex.map(deflate_piece, starts))
            with open(gzp, "wb") as f:
                f.write(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03")
                for part in parts:
                    f.write(part)
                f.write(int(zlib.crc32(host) & 0xFFFFFFFF).to_bytes(4, "little") + int(nbytes & 0xFFFFFFFF).to_bytes(4, "little"))
            del parts
            blob = bgzf.compress(host, 6, threads=cores)
            with open(bgz, "wb") as f:
                f.write(blob)

            def run(parser, path):
                st_ = parser.stream_open(path, capi.SOURCE_AUTO, region_mib << 20)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = 0
                while True:
                    res, off, first = parser.stream_next_result(st_, capi.WANT_BATCHES)
                    n += int(res.n_records)
                    if res.stop.code != capi.OK:
                        assert res.stop.code == capi.EOF, res.stop.text
                        break
                torch.cuda.synchronize()
                wall = time.perf_counter() - t0
                s_ = parser.stream_stats(st_)
                parser.stream_close(st_)
                assert n == reads, (n, reads)
                return {"wall_s": wall, "reads_per_s": reads / wall, "uncompressed_gb_per_s": nbytes / wall / 1e9,
                        "reader_busy_s": s_.reader_busy_s, "gpu_pass_s": s_.parse_s, "caller_wait_reader_s": s_.wait_reader_s,
                        "regions": int(s_.regions), "h2d_compressed_s": s_.h2d_s, "inflate_kernels_s": s_.inflate_s,
                        "compressed_bytes_over_pcie": int(s_.compressed_bytes), "launch_s": s_.launch_s,
                        "wait_inflate_s": s_.wait_inflate_s}
            run(g, plain)                     # warm the page cache and the arenas
            run(g, bgz)
            out = {"bgzf_device_inflate": run(g, bgz), "bgzf_host_threads": run(gh, bgz), "plain_file": run(g, plain)}
            # ordinary gzip: decoded speculatively in parallel by every host thread (bsq_pgzip.h), and by one zlib thread
            out["gzip_parallel_host_threads"] = run(g, gzp)
            if gib <= 1.0:                                                             # 0.26 GB/s: only on small inputs
                g1 = B.GpuParser(False, False, schema, 4096, device_id=local, inflate_threads=1)
                out["gzip_zlib_reader_thread"] = run(g1, gzp)
                g1.close()
            else:
                out["gzip_zlib_reader_thread"] = None
            # CPU arm: zlib over the same BGZF members on every host thread (what the reference's parallel reader does)
            offs, pos = [], 0
            while pos < len(blob):
                total = (blob[pos + 16] | (blob[pos + 17] << 8)) + 1
                offs.append((pos, total))
                pos += total
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                def work(k):
                    tot = 0
                    for a, t_ in offs[k::cores]:
                        tot += len(zlib.decompress(blob[a + 18:a + t_ - 8], -15))
                    return tot
                assert sum(ex.map(work, range(cores))) == nbytes
            out["cpu_zlib_all_threads_inflate_only"] = {"wall_s": time.perf_counter() - t0, "threads": cores,
                                                        "uncompressed_gb_per_s": nbytes / (time.perf_counter() - t0) / 1e9}
            out.update({"reads": reads, "uncompressed_bytes": nbytes, "bgzf_bytes": len(blob), "gzip_bytes": os.path.getsize(gzp),
                        "region_mib": region_mib})
            return out
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
            g.close(); gh.close()

    sampler = ClockSampler(local)
    sampler.start()

    if args.gzip:
        gib = args.gib if args.gib != 10.0 else 4.0
        leg = gzip_leg(gib, args.region_mib)
        clocks = sampler.stop()
        best = leg["bgzf_device_inflate"]
        if rank == 0:
            emit(json.dumps({
                "metric": "fastq_reads_per_s", "value": best["reads_per_s"], "unit": "reads/s", "n_gpus": 1, "steps": 1, "warmup": 1,
                "ms_per_step": best["wall_s"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "parsed_gb_per_s": best["uncompressed_gb_per_s"],
                "config": {"workload": "configs[4]: %.2f GiB (uncompressed) 150 bp FASTQ as BGZF (gzip -6, 64 KiB members), file -> "
                                       "bsq_stream_next(batches 4096): compressed members over PCIe, inflated on the device, "
                                       "regions of %d MiB" % (leg["uncompressed_bytes"] / GIB, leg["region_mib"])},
                "roofline": None, "cpu_baseline": {"value": leg["cpu_zlib_all_threads_inflate_only"]["uncompressed_gb_per_s"], "unit": "GB/s",
                                                   "cores": leg["cpu_zlib_all_threads_inflate_only"]["threads"], "kind": "port",
                                                   "sample": "zlib over the same BGZF members on every host thread, inflate only"},
                "e2e": {"value": best["reads_per_s"], "unit": "reads/s", "h2d_bytes_per_step": leg["bgzf_bytes"], "d2h_bytes_per_step": 0,
                        "result": "DeviceFastqBatch SoA per region"},
                "gzip": leg, "clocks": clocks}))
        return

    if args.shard_stream:
        mn, mx = (75, 300) if args.mixed else (args.read_len, args.read_len)
        leg = shard_stream_leg(args.gib, mn, mx, args.steps, max(args.warmup, 3), args.mode)
        clocks = sampler.stop()
        resolve_ms = leg["ms"][2]
        algo_bytes = leg["algo"] * leg["own_reads"]
        roofline = {"bound": "hbm", "kernel": "k_resolve", "pass": "shard summary + two-pass parse of the own records",
                    "achieved": algo_bytes / (resolve_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": algo_bytes / (resolve_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_record": leg["algo"], "records_per_launch": leg["own_reads"] / leg["n_windows"],
                    "launches_per_step": leg["n_windows"], "avg_launch_ms": resolve_ms / leg["n_windows"],
                    "summarize_ms_per_step": leg["ms"][0], "tail_rebase_ms_per_step": leg["ms"][3], "step_device_ms": leg["dev_ms"],
                    "note": "rank 0's kernels; `value` is the whole job (max over ranks, summaries + collectives inside)"}
        if rank == 0:
            emit(json.dumps({
                "metric": "fastq_reads_per_s", "value": leg["value"], "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "parsed_gb_per_s": leg["bytes"] / (leg["ms_per_step"] * 1e-3) / 1e9,
                "config": workload_config(args, world, leg["reads"], leg["bytes"], leg["record_bytes"]),
                "roofline": roofline, "cpu_baseline": None, "e2e": None, "gpu_launches": leg["launches"], "clocks": clocks,
                "timing": "wall clock between barrier+synchronize pairs (max over ranks)"}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------------
    # the headline: this rank's shard of an (N * M)-record stream, generated on the device
    # ------------------------------------------------------------------------------------------------
    schema = B.parse_schema("sanger" if args.validate else "illumina_1.8")
    mn, mx = (75, 300) if args.mixed else (args.read_len, args.read_len)
    M = L.bsq_compute_num_reads_for_size(int(args.gib * GIB), mn, mx)
    total_reads = M * world
    stream_reads = total_reads              # ids are zero padded to the width of the stream's last index
    if args.id_digits:
        stream_reads = max(total_reads, 10 ** (args.id_digits - 1) + 1)
    st = Stream(stream_reads, mn, mx)
    gpu = B.GpuParser(args.validate, args.validate, schema, 4096, device_id=local)
    lo_off, hi_off = st.offset(rank * M), st.offset((rank + 1) * M)
    size = hi_off - lo_off
    rec_bytes = size / M                       # average bytes per record
    bases_expected = st.bases(rank * M, (rank + 1) * M)
    buf = torch.empty(size + 256, dtype=torch.uint8, device=dev)
    assert gpu.synth_device(buf.data_ptr(), size, stream_reads, rank * M, M, mn, mx, 2, 40, schema) == size
    if args.crlf:
        # "\n" -> "\r\n" on the device: the id loses its "\r" (utils.mojo:221-242), sequence and quality keep it
        nl = buf[:size] == 10
        pos = torch.arange(size, device=dev, dtype=torch.int64) + torch.cumsum(nl, 0)
        out = torch.empty(size + 4 * M + 256, dtype=torch.uint8, device=dev)
        out[pos] = buf[:size]
        out[pos[nl] - 1] = 13
        del nl, pos
        buf, size = out, size + 4 * M
        bases_expected += M
        rec_bytes = size / M
        lo_off = 0
    want = capi.WANT_BATCHES if args.mode == "batches" else capi.WANT_OFFSETS
    # algorithmic bytes per record: R read + (2L + I + 16) written, or R + 20 for views
    algo = rec_bytes + (2 * (bases_expected / M) + st.id_len + 16 if args.mode == "batches" else 20)

    def step():
        r = gpu.parse_device(buf.data_ptr(), size, lo_off, rank * M, True, want)
        assert r.n_records == M and r.stop.code == capi.EOF, (r.n_records, r.stop.text)
        return r

    # nvidia-smi needs ~100 ms to start: it was launched above so that it is sampling (every 20 ms) while the
    # timed steps run; warm-up samples are under the same load
    for _ in range(max(args.warmup, 3)):
        res = step()
    assert res.n_bases == bases_expected

    # ---- the SoA of the whole pass, checked on the device (untimed): every record, not only the counts ------
    soa_check = None
    if args.mode == "batches":
        v = gpu.soa_view()
        seq = torch.as_tensor(_DevPtr(v.sequence_buffer, v.sequence_bytes), device=dev)
        qual = torch.as_tensor(_DevPtr(v.qual_buffer, v.seq_len), device=dev)
        ids = torch.as_tensor(_DevPtr(v.id_buffer, v.total_id_bytes), device=dev)
        ends = torch.as_tensor(_DevPtr(v.ends, M, "<i8"), device=dev)
        id_ends = torch.as_tensor(_DevPtr(v.id_ends, M, "<i8"), device=dev)
        ok = int(v.seq_len) == bases_expected and int(v.total_id_bytes) == M * st.id_len and int(v.num_records) == M
        if mn == mx and not args.crlf:
            # constant stride: the arenas are strided gathers of the input, the ends arithmetic progressions per batch
            rec, hdr = st.overhead + 2 * mn, st.id_len + 2
            recs = buf[:size].view(M, rec)
            ok = ok and bool(torch.equal(seq.view(M, mn), recs[:, hdr:hdr + mn]))
            ok = ok and bool(torch.equal(qual.view(M, mn), recs[:, hdr + mn + 3:hdr + 2 * mn + 3]))
            ok = ok and bool(torch.equal(ids.view(M, st.id_len), recs[:, 1:1 + st.id_len]))
            k = torch.arange(M, device=dev)
            ok = ok and bool(torch.equal(ends, (k % 4096 + 1) * mn)) and bool(torch.equal(id_ends, (k % 4096 + 1) * st.id_len))
            del k, recs
            soa_check = "all %d records: sequence / quality / id arenas == strided gather of the input, ends exact" % M
        else:
            # every input byte is an id, sequence or quality byte, '@', '+' or a line end (CRLF: the '\r' of the id
            # and '+' lines is dropped, the others stay in the arenas): a linear checksum over the arenas
            total = int(buf[:size].sum(dtype=torch.int64))
            got = int(seq.sum(dtype=torch.int64)) + int(qual.sum(dtype=torch.int64)) + int(ids.sum(dtype=torch.int64))
            fixed = M * (ord("@") + ord("+") + 4 * 10) + (M * 13 * 2 if args.crlf else 0)
            ok = ok and got + fixed == total
            ok = ok and bool(torch.all(ends[1:] - ends[:-1] != 0)) and int(id_ends[-1]) == (M - (M - 1) // 4096 * 4096) * st.id_len
            soa_check = "byte-sum checksum of the three arenas == input minus delimiters; sizes and id ends exact"
        assert ok, "SoA check failed"
        del seq, qual, ids, ends, id_ends

    # ---- timed region ------------------------------------------------------------------------------
    wall_max, ms_avg, dev_ms_max, launches = timed(step, args.steps, gpu)
    clocks = sampler.stop()
    dev_ms = ms_avg[4]
    # the one collective of the path: total reads / bases
    reads, bases = M, bases_expected
    if dist is not None:
        reads, bases = sharding.allreduce_counts(dist, M, bases_expected, device=dev)
    assert reads == total_reads
    ms_per_step = wall_max / args.steps * 1e3
    value = total_reads / (wall_max / args.steps)

    # ---- roofline of the dominant kernel (k_resolve: read R, write the SoA) ----------------------
    n_windows = int(res.n_windows)
    resolve_ms = ms_avg[2]                        # all k_resolve launches of one step
    achieved = algo * M / (resolve_ms * 1e-3) / 1e9
    traffic = None
    try:
        # measured DRAM bytes per record of k_resolve (one ncu --set full capture, profiles/traffic.json),
        # scaled to the records one launch of this run processes
        per_rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_resolve_bytes_per_record"]
        traffic = per_rec * M / n_windows if args.mode == "batches" and not args.validate and mn == mx == 150 else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_resolve", "pass": "two-pass (k_summarize + k_scan_runs, then k_resolve)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_record": algo, "records_per_launch": M / n_windows,
                "launches_per_step": n_windows, "avg_launch_ms": resolve_ms / n_windows,
                "summarize_ms_per_step": ms_avg[0], "tail_rebase_ms_per_step": ms_avg[3],
                "step_device_ms": dev_ms,
                "step_frac": algo * M / (dev_ms * 1e-3) / 1e9 / peak}

    # ---- sub-results: the other BASELINE configs on the same box, same run (short) ----------------
    sub = None
    plain = not (args.validate or args.mixed or args.crlf or args.mode != "batches" or args.id_digits or args.read_len != 150)
    if plain and not args.no_sub:
        sub = {}
        k = max(1, args.sub_steps)

        def step_views():   # views(): offsets only
            r = gpu.parse_device(buf.data_ptr(), size, lo_off, rank * M, True, capi.WANT_OFFSETS)
            assert r.n_records == M and r.stop.code == capi.EOF
        step_views()
        w, ms, _, _ = timed(step_views, k, gpu)
        sub["views"] = {"workload": "configs[1] input, views(): offsets table only, per GPU", "value": total_reads / (w / k),
                        "unit": "reads/s", "ms_per_step": w / k * 1e3, "steps": k,
                        "k_resolve_ms_per_launch": ms[2] / n_windows, "summarize_ms_per_step": ms[0]}
        # configs[2]: validation on (sanger == illumina_1.8 numerically: the same bytes)
        gv = B.GpuParser(True, True, B.parse_schema("sanger"), 4096, device_id=local)

        def step_val():
            r = gv.parse_device(buf.data_ptr(), size, lo_off, rank * M, True, capi.WANT_BATCHES)
            assert r.n_records == M and r.stop.code == capi.EOF and r.n_bases == bases_expected
        step_val()
        w, ms, _, _ = timed(step_val, k, gv)
        sub["configs[2]"] = {"workload": "configs[2]: the same 10 GiB, check_ascii + check_quality, sanger, batches(4096), per GPU",
                             "value": total_reads / (w / k), "unit": "reads/s", "ms_per_step": w / k * 1e3, "steps": k,
                             "k_resolve_ms_per_launch": ms[2] / n_windows, "summarize_ms_per_step": ms[0]}
        gv.close()
        # configs[3]: one mixed-length 10 GiB stream cut over the ranks at arbitrary byte offsets
        leg = shard_stream_leg(args.gib, 75, 300, k, 1)
        sub["configs[3]"] = {"workload": f"configs[3]: ONE {args.gib:g} GiB mixed read-length (75-300 bp) stream, validation OFF, "
                                         f"batches(4096), {world} byte shard(s) cut at arbitrary offsets (summary -> all-gather -> "
                                         "shard prefix -> own records + halo -> all-reduce)", "scaling": "strong",
                             "value": leg["value"], "unit": "reads/s", "ms_per_step": leg["ms_per_step"], "steps": k,
                             "reads_total": leg["reads"], "parsed_gb_per_s": leg["bytes"] / (leg["ms_per_step"] * 1e-3) / 1e9,
                             "k_resolve_ms_per_launch": leg["ms"][2] / leg["n_windows"], "summarize_ms_per_step": leg["ms"][0]}
        if rank == 0 and world == 1:
            # configs[4] (scaled to 1 GiB to keep the default run short; `bench.py --gzip` runs the 4 GiB case)
            try:
                gl = gzip_leg(1.0)
                sub["configs[4]"] = {"workload": "configs[4] scaled: 1 GiB (uncompressed) 150 bp FASTQ as BGZF / gzip files -> "
                                                 "bsq_stream_next(batches 4096)", "unit": "GB/s of FASTQ text",
                                     "bgzf_device_inflate": gl["bgzf_device_inflate"]["uncompressed_gb_per_s"],
                                     "bgzf_host_threads": gl["bgzf_host_threads"]["uncompressed_gb_per_s"],
                                     "gzip_parallel_host_threads": gl["gzip_parallel_host_threads"]["uncompressed_gb_per_s"],
                                     "gzip_zlib_reader_thread": (gl["gzip_zlib_reader_thread"] or {}).get("uncompressed_gb_per_s"),
                                     "plain_file": gl["plain_file"]["uncompressed_gb_per_s"],
                                     "cpu_zlib_all_threads_inflate_only": gl["cpu_zlib_all_threads_inflate_only"]["uncompressed_gb_per_s"],
                                     "value": gl["bgzf_device_inflate"]["reads_per_s"], "ms_per_step": gl["bgzf_device_inflate"]["wall_s"] * 1e3}
            except Exception as e:   # (no room in /dev/shm, ...): the headline does not depend on it
                sub["configs[4]"] = {"skipped": repr(e)[:200]}

    # ---- e2e: the same pass starting from pinned host memory (H2D inside the timed region) -------
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        except Exception:
            host = torch.empty(size, dtype=torch.uint8)
        host.copy_(buf[:size])
        harr = host.numpy()
        r = gpu.parse_host(harr, lo_off, rank * M, True, want)   # warm: allocates the device staging copy
        assert r.n_records == M
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = gpu.parse_host(harr, lo_off, rank * M, True, want)
            assert r.n_records == M and r.stop.code == capi.EOF
        barrier()
        (dt,) = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        d2h = 8 * (int(r.n_batches) + 1) * 2 + 8 + 16 + 168   # batch directory + error word + scan totals
        e2e = {"value": total_reads / dt, "unit": "reads/s", "h2d_bytes_per_step": size,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
               "result": "DeviceFastqBatch SoA left on the device + host batch directory", "host": host_info}
        def all_ranks_ok(ok: bool) -> bool:
            """a leg with collectives inside runs only when every rank got its buffers (no rank may wait at a barrier alone)"""
            (bad,) = max_over_ranks(0.0 if ok else 1.0)
            return bad == 0.0

        outs = None
        if args.mode == "batches":
            # ... and with the reference's HOST product: the whole FastqBatch SoA copied back to pinned memory
            v = gpu.soa_view()
            try:
                outs = [torch.empty(int(nb), dtype=dt_, pin_memory=True) for nb, dt_ in (
                    (v.sequence_bytes, torch.uint8), (v.seq_len, torch.uint8), (v.total_id_bytes, torch.uint8),
                    (M, torch.int64), (M, torch.int64))]
            except Exception:
                outs = None
            if not all_ranks_ok(outs is not None):
                outs = None
                e2e["host_batch"] = {"skipped": "no pinned host memory for the SoA arrays on some rank"}
        if outs is not None:
            gpu.soa_to_host(*outs)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                r = gpu.parse_host(harr, lo_off, rank * M, True, want)
                gpu.soa_to_host(*outs)
            barrier()
            (dth,) = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
            back = int(v.sequence_bytes) + int(v.seq_len) + int(v.total_id_bytes) + 16 * M
            assert int(outs[3][-1]) == int(torch.as_tensor(_DevPtr(v.ends, M, "<i8"), device=dev)[-1])
            e2e["host_batch"] = {"value": total_reads / dth, "unit": "reads/s", "ms_per_step": dth * 1e3,
                                 "h2d_bytes_per_step": size, "d2h_bytes_per_step": back + d2h,
                                 "result": "host FastqBatch SoA (five arrays, pinned) via bsq_soa_to_host after the pass"}
            # ... and the same product with both PCIe directions busy: two parser handles alternate regions, the SoA of
            # region k travels back while region k+1 travels in (blazeseq_b200/pipeline.py)
            ref_ends = torch.as_tensor(_DevPtr(v.ends, M, "<i8"), device=dev).cpu()
            tried, why = {}, None
            for region in (256 << 20, 1 << 30):
                pipe = None
                try:
                    pipe = B.HostBatchPipeline(lambda: B.GpuParser(args.validate, args.validate, schema, 4096, device_id=local),
                                               region_bytes=region)
                    outs[3].zero_()
                    got = pipe.run(harr, *outs, stream_offset=lo_off, first_record=rank * M)     # warm-up: arenas, staging
                    assert got[0] == M and pipe.stop.code == capi.EOF, (got, pipe.stop.text)
                    assert torch.equal(outs[3], ref_ends), "pipelined ends differ from the one-pass SoA"
                    ok = True
                except Exception as e:   # the headline does not depend on it
                    ok, why = False, repr(e)[:300]
                if all_ranks_ok(ok):
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(args.e2e_steps):
                        got = pipe.run(harr, *outs, stream_offset=lo_off, first_record=rank * M)
                    barrier()
                    (dtp,) = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
                    assert got == (M, int(v.sequence_bytes), int(v.seq_len), int(v.total_id_bytes)), got
                    tried[region] = dtp
                if pipe is not None:
                    pipe.close()
            if tried:
                region, dtp = min(tried.items(), key=lambda kv: kv[1])
                e2e["host_batch_pipelined"] = {
                    "value": total_reads / dtp, "unit": "reads/s", "ms_per_step": dtp * 1e3, "h2d_bytes_per_step": size,
                    "d2h_bytes_per_step": back, "region_bytes": region,
                    "ms_per_step_by_region_mib": {str(r >> 20): t * 1e3 for r, t in tried.items()},
                    "result": "host FastqBatch SoA (five arrays, pinned); two parser handles alternate regions so that D2H of "
                              "region k overlaps H2D + passes of region k+1 (HostBatchPipeline)"}
            else:
                e2e["host_batch_pipelined"] = {"skipped": why or "failed on another rank"}
            del outs
        del host, harr

    # ---- CPU baseline beside it (rank 0, N=1): the oracle port on a bounded sample ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        cores = os.cpu_count() or 1
        sample_reads = min(M, int(args.cpu_sample_gib * GIB / rec_bytes))
        sample_bytes = (st.offset(rank * M + sample_reads) - st.offset(rank * M)) + (4 * sample_reads if args.crlf else 0)
        sample = buf[:sample_bytes].cpu().numpy()
        cfg = O.config(args.validate, args.validate, "sanger" if args.validate else "illumina_1.8")
        mode = 1 if args.mode == "batches" else 0

        def run(threads, budget):
            O.baseline_mt(sample, cfg, mode, 4096, threads)
            n_it, t0 = 0, time.perf_counter()
            while True:
                n, b, code = O.baseline_mt(sample, cfg, mode, 4096, threads)
                assert n == sample_reads and code == 0
                n_it += 1
                if time.perf_counter() - t0 > budget:
                    break
            return sample_reads * n_it / (time.perf_counter() - t0)
        one = run(1, 6.0)
        allc = run(cores, 6.0)
        cpu = {"value": allc, "unit": "reads/s", "cores": cores, "kind": "port", "value_1core": one,
               "sample": f"first {sample_reads} reads ({sample.size / GIB:.2f} GiB) of this workload, {args.mode}"}

    if rank == 0:
        out = {
            "metric": "fastq_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "parsed_gb_per_s": size * world / (wall_max / args.steps) / 1e9,
            "config": workload_config(args, world, M, size, rec_bytes),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "soa_check": soa_check, "sub_results": sub,
            "timing": "wall clock between barrier+synchronize pairs (max over ranks); kernels timed with CUDA events "
                      "on the parser's stream",
        }
        emit(json.dumps(out))
    gpu.close()
    if dist is not None:
        dist.destroy_process_group()


def emit(line: str) -> None:
    """The one JSON line goes to the process's real stdout; everything else that lands on fd 1 while the
    bench runs (NCCL's version banner, library chatter) was redirected to stderr in __main__."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
