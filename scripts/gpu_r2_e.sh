# round 2: full parity suite on the shipped library + bench lines (LF / 320-byte stride / mixed / validate / views / CRLF)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
show() { python - "$@" <<'P'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "stride %.0f k_resolve %.4f ms frac %.3f summarize %.3f tail %.3f step %.3f ms %.2f Greads/s" % (d["config"]["record_bytes"], r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], r["tail_rebase_ms_per_step"], d["ms_per_step"], d["value"] / 1e9))
    except Exception as e:
        print(f, "failed", e)
P
}
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 $B > gpurun_out/e_main.json 2> gpurun_out/e.err
timeout 300 $B --id-digits 9 > gpurun_out/e_main_320.json 2>> gpurun_out/e.err
timeout 300 $B --mixed > gpurun_out/e_main_mixed.json 2>> gpurun_out/e.err
timeout 300 $B --validate > gpurun_out/e_main_validate.json 2>> gpurun_out/e.err
timeout 300 $B --mode views > gpurun_out/e_main_views.json 2>> gpurun_out/e.err
timeout 300 $B --gib 4 > gpurun_out/e_lf4.json 2>> gpurun_out/e.err
timeout 300 $B --gib 4 --crlf > gpurun_out/e_crlf4.json 2>> gpurun_out/e.err
show gpurun_out/e_main.json gpurun_out/e_main_320.json gpurun_out/e_main_mixed.json gpurun_out/e_main_validate.json gpurun_out/e_main_views.json gpurun_out/e_lf4.json gpurun_out/e_crlf4.json
tail -5 gpurun_out/e.err
