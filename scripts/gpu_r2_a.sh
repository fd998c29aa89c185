# round 2, first GPU visit: parity of the direct-copy k_resolve, bench line, stride sweep
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; cat gpurun_out/a_bench.json | cut -c1-1800
bash scripts/gpu_stride_sweep.sh
# tuning variants (scripts/build_variants.sh): CTAs per SM, a second tile buffer
V=blazeseq_b200/lib/variants
for tag in c5 c7 c8 s2c5; do
  [ -f $V/lib_$tag.so ] || continue
  BSQ_LIB=$V/lib_$tag.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/a_var_$tag.json 2>> gpurun_out/a_bench.err
  BSQ_LIB=$V/lib_$tag.so timeout 300 python bench.py --gib 2 --steps 5 --warmup 3 --no-cpu --no-e2e --id-digits 9 > gpurun_out/a_var_${tag}_320.json 2>> gpurun_out/a_bench.err
  python - $tag <<'P'
import json, sys
for suf in ("", "_320"):
    try:
        d = json.load(open("gpurun_out/a_var_%s%s.json" % (sys.argv[1], suf))); r = d["roofline"]
        print(sys.argv[1] + suf, "k_resolve %.4f ms frac %.3f summarize %.3f step %.3f ms %.2f Greads/s" % (r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], d["ms_per_step"], d["value"] / 1e9))
    except Exception as e:
        print(sys.argv[1] + suf, "failed", e)
P
done
tail -5 gpurun_out/a_bench.err
