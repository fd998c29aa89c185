#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200.

Metric (BASELINE.json): FASTQ reads/s (and parsed GB/s) of the scan / validate / FastqBatch-pack hot
path.  Workload at every N: BASELINE.json configs[1] -- 10 GiB in-memory 150 bp Illumina FASTQ
(33,659,618 reads, generate_synthetic_fastq_buffer semantics), schema illumina_1.8, validation OFF,
batches(4096) -- PER GPU (weak scaling: rank r holds records [r*M, (r+1)*M) of an N*M-record stream).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

One "step" = one pass of the hot path over the whole per-GPU input, already resident in HBM
(`value`), or starting from pinned HOST memory through bsq_parse_host (`e2e`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
ALGO_BYTES_BATCHES = 648   # R=319 read + 2L+I+16 = 329 written, per 150 bp record
ALGO_BYTES_VIEWS = 339


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
    all host threads, on a bounded sample of the same workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py as O
    from concurrent.futures import ThreadPoolExecutor

    O.build()
    cores = os.cpu_count() or 1
    n_total = O.compute_num_reads_for_size(10 * GIB, 150, 150)
    sample_reads = min(n_total, int(args.ref_sample_gib * GIB) // 319)
    rec = 319
    data = np.empty(sample_reads * rec, np.uint8)
    parts = max(1, min(cores, 32))
    step = (sample_reads + parts - 1) // parts

    def gen(i):
        first = i * step
        cnt = min(step, sample_reads - first)
        if cnt > 0:
            O.synth(n_total, 150, 150, 2, 40, "illumina_1.8", first=first, count=cnt,
                    out=data[first * rec:(first + cnt) * rec])
    with ThreadPoolExecutor(parts) as ex:
        list(ex.map(gen, range(parts)))
    cfg = O.config(False, False, "illumina_1.8")
    for _ in range(args.warmup):
        O.baseline_mt(data, cfg, 1, 4096, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, bases, code = O.baseline_mt(data, cfg, 1, 4096, cores)
        assert n == sample_reads and code == 0
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_reads / dt
    # the reference's FastqParser is a sequential object (one parser = one thread, record.mojo:439): the same
    # sample through ONE thread is what a single BlazeSeq parser delivers; `value` shards the sample by
    # newline rank over every host thread, which the reference itself does not do
    t1 = time.perf_counter()
    O.baseline_mt(data, cfg, 1, 4096, 1)
    one_thread = sample_reads / (time.perf_counter() - t1)
    sample = f"first {sample_reads} reads ({data.size / GIB:.2f} GiB) of the 10 GiB 150 bp stream, batches(4096)"
    emit(json.dumps({
        "impl": "reference", "metric": "fastq_reads_per_s", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "parsed_gb_per_s": data.size / dt / 1e9,
        "config": {"workload": "configs[1]: 10 GiB in-memory 150 bp Illumina FASTQ, illumina_1.8, validation OFF, "
                               "batches(4096)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample,
                         "value_1core": one_thread},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
    ap.add_argument("--ref-sample-gib", type=float, default=1.0)
    ap.add_argument("--cpu-sample-gib", type=float, default=1.0)
    ap.add_argument("--mode", default="batches", choices=["batches", "views"])
    ap.add_argument("--validate", action="store_true", help="configs[2]: check_ascii + check_quality, sanger")
    ap.add_argument("--mixed", action="store_true", help="configs[3]: mixed read length 75-300 bp instead of 150 bp")
    ap.add_argument("--read-len", type=int, default=150, help="read length of the synthetic stream (record stride sweeps)")
    ap.add_argument("--id-digits", type=int, default=0, help="zero-padded id width (0: that of the stream's read count)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- input: this rank's shard of an (N * M)-record stream, generated on the device --------
    schema = B.parse_schema("sanger" if args.validate else "illumina_1.8")
    mn, mx = (75, 300) if args.mixed else (args.read_len, args.read_len)
    M = capi.lib().bsq_compute_num_reads_for_size(int(args.gib * GIB), mn, mx)
    total_reads = M * world
    stream_reads = total_reads              # ids are zero padded to the width of the stream's last index
    if args.id_digits:
        stream_reads = max(total_reads, 10 ** (args.id_digits - 1) + 1)
    digits = len(str(stream_reads - 1))
    gpu = B.GpuParser(args.validate, args.validate, schema, 4096, device_id=local)
    # byte range of this rank's records inside the (world * M)-record stream (utils.mojo:753-768:
    # len_i = mn + (31 i + 7) % (mx - mn + 1), header "@read_<i zero padded to `digits`>\n")
    L = capi.lib()
    period = mx - mn + 1
    pref = [0]
    for j in range(period):
        pref.append(pref[-1] + (31 * j + 7) % period)

    def stream_offset(i):
        lens = i * mn + (i // period) * pref[period] + pref[i % period]
        return i * (6 + digits + 1 + 4) + 2 * lens
    lo_off, hi_off = stream_offset(rank * M), stream_offset((rank + 1) * M)
    size = hi_off - lo_off
    rec_bytes = size / M                       # average bytes per record
    bases_expected = (size - M * (6 + digits + 1 + 4)) // 2
    buf = torch.empty(size + 256, dtype=torch.uint8, device=dev)
    assert gpu.synth_device(buf.data_ptr(), size, stream_reads, rank * M, M, mn, mx, 2, 40, schema) == size
    want = capi.WANT_BATCHES if args.mode == "batches" else capi.WANT_OFFSETS
    # algorithmic bytes per record: R read + (2L + I + 16) written, or R + 20 for views
    id_len = 5 + digits
    algo = rec_bytes + (2 * (bases_expected / M) + id_len + 16 if args.mode == "batches" else 20)

    def step():
        r = gpu.parse_device(buf.data_ptr(), size, lo_off, rank * M, True, want)
        assert r.n_records == M and r.stop.code == capi.EOF, (r.n_records, r.stop.text)
        return r

    # nvidia-smi needs ~100 ms to start: launch it before the warm-up so that it is sampling (every
    # 100 ms) while the timed steps run; warm-up samples are under the same load
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = step()
    assert res.n_bases == bases_expected

    # ---- timed region ------------------------------------------------------------------------------
    barrier()
    ms_sum = [0.0] * 5
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
        ms, nl = gpu.timing()
        ms_sum = [a + b for a, b in zip(ms_sum, ms)]
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = ms_sum[4] / args.steps
    t = torch.tensor([wall, dev_ms, float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, dev_ms_max, launches = float(t[0]), float(t[1]), int(t[2])
    # the one collective of the path: total reads / bases
    reads, bases = M, bases_expected
    if dist is not None:
        reads, bases = sharding.allreduce_counts(dist, M, bases_expected, device=dev)
    assert reads == total_reads
    ms_per_step = wall_max / args.steps * 1e3
    value = total_reads / (wall_max / args.steps)

    # ---- roofline of the dominant kernel (k_resolve: read R, write the SoA) ----------------------
    peak, peak_src = measured_peaks()
    n_windows = int(res.n_windows)
    resolve_ms = ms_sum[2] / args.steps           # all k_resolve launches of one step
    achieved = algo * M / (resolve_ms * 1e-3) / 1e9
    traffic = None
    try:
        # measured DRAM bytes per record of k_resolve (one ncu --set full capture, profiles/traffic.json),
        # scaled to the records one launch of this run processes
        per_rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_resolve_bytes_per_record"]
        traffic = per_rec * M / n_windows if args.mode == "batches" and not args.validate else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_resolve", "pass": "two-pass (k_summarize + k_scan_runs, then k_resolve)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_record": algo, "records_per_launch": M / n_windows,
                "launches_per_step": n_windows, "avg_launch_ms": resolve_ms / n_windows,
                "summarize_ms_per_step": ms_sum[0] / args.steps, "tail_rebase_ms_per_step": ms_sum[3] / args.steps,
                "step_device_ms": dev_ms,
                "step_frac": algo * M / (dev_ms * 1e-3) / 1e9 / peak}

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
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = 8 * (int(r.n_batches) + 1) * 2 + 8 + 16 + 168   # batch directory + error word + scan totals
        e2e = {"value": total_reads / float(tt[0]), "unit": "reads/s", "h2d_bytes_per_step": size,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(tt[0]) * 1e3,
               "result": "DeviceFastqBatch SoA left on the device + host batch directory"}
        del host, harr

    # ---- CPU baseline beside it (rank 0, N=1): the oracle port on a bounded sample ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        cores = os.cpu_count() or 1
        sample_reads = min(M, int(args.cpu_sample_gib * GIB / rec_bytes))
        sample_bytes = stream_offset(rank * M + sample_reads) - lo_off
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
            "config": {"workload": ("configs[3]: 10 GiB mixed read-length (75-300 bp) FASTQ, validation OFF" if args.mixed
                                    else "configs[2]: 10 GiB 150 bp FASTQ, check_ascii+check_quality, sanger" if args.validate
                                    else "configs[1]: 10 GiB in-memory 150 bp Illumina FASTQ, illumina_1.8, validation OFF")
                       + f", {args.mode}(4096), per GPU", "reads_per_gpu": M, "bytes_per_gpu": size,
                       "record_bytes": rec_bytes, "l2": "input (>=10 GB) is larger than L2; no flush needed",
                       "parallelism": f"{world} x record-aligned shard, NCCL all-reduce of counts only"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
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
