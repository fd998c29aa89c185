# round 2: bank-occupancy test in front of the start rotation, k_summarize trims and occupancy variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "stride or random or screen or corpus or degenerate or all_newlines or long_reads or scale or config1 or mixed" > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
show() { python - "$@" <<'P'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "stride %.0f k_resolve %.4f ms frac %.3f summarize %.3f step %.3f ms %.2f Greads/s" % (d["config"]["record_bytes"], r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], d["ms_per_step"], d["value"] / 1e9))
    except Exception as e:
        print(f, "failed", e)
P
}
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
V=blazeseq_b200/lib/variants
timeout 300 $B > gpurun_out/d_main.json 2> gpurun_out/d.err
timeout 300 $B --id-digits 9 > gpurun_out/d_main_320.json 2>> gpurun_out/d.err
timeout 300 $B --mixed > gpurun_out/d_main_mixed.json 2>> gpurun_out/d.err
timeout 300 $B --validate > gpurun_out/d_main_validate.json 2>> gpurun_out/d.err
timeout 300 $B --mode views > gpurun_out/d_main_views.json 2>> gpurun_out/d.err
show gpurun_out/d_main.json gpurun_out/d_main_320.json gpurun_out/d_main_mixed.json gpurun_out/d_main_validate.json gpurun_out/d_main_views.json
for tag in s1c8 s1c10 s2c5; do
  [ -f $V/lib_$tag.so ] || continue
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B > gpurun_out/d_$tag.json 2>> gpurun_out/d.err
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B --mode views > gpurun_out/d_${tag}_views.json 2>> gpurun_out/d.err
  show gpurun_out/d_$tag.json gpurun_out/d_${tag}_views.json
done
tail -3 gpurun_out/d.err
