# round 2: full parity suite + the complete default bench line (sub-results, e2e, cpu baseline) + reference arm + CRLF
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
show() { python - "$@" <<'P'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "stride %.0f k_resolve %.4f ms frac %.3f summarize %.3f tail %.3f step %.3f ms %.2f Greads/s step_frac %s" % (d["config"]["record_bytes"], r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], r["tail_rebase_ms_per_step"], d["ms_per_step"], d["value"] / 1e9, r.get("step_frac")))
        for k, v in (d.get("sub_results") or {}).items(): print("   sub", k, "%.2f Greads/s %.3f ms" % (v["value"] / 1e9, v["ms_per_step"]))
        e = d.get("e2e")
        if e: print("   e2e %.1f Mreads/s %.1f ms; host_batch %s" % (e["value"] / 1e6, e["ms_per_step"], (e.get("host_batch") or {}).get("ms_per_step")))
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"] / 1e6, d["cpu_baseline"]["cores"])
    except Exception as e:
        print(f, "failed", e)
P
}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f.err; show gpurun_out/f_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/f_bench_reference.json 2>> gpurun_out/f.err; cut -c1-400 gpurun_out/f_bench_reference.json
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sub"
timeout 300 $B --id-digits 9 > gpurun_out/f_320.json 2>> gpurun_out/f.err
timeout 300 $B --gib 4 > gpurun_out/f_lf4.json 2>> gpurun_out/f.err
timeout 300 $B --gib 4 --crlf > gpurun_out/f_crlf4.json 2>> gpurun_out/f.err
timeout 300 $B --mixed --shard-stream > gpurun_out/f_shard1.json 2>> gpurun_out/f.err
show gpurun_out/f_320.json gpurun_out/f_lf4.json gpurun_out/f_crlf4.json gpurun_out/f_shard1.json
tail -5 gpurun_out/f.err
