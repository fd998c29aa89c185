# round 2: staged copy with per-bank start rotation (shipped lib) vs the direct copy (variants), same box
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
show() { python - "$@" <<'P'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "k_resolve %.4f ms frac %.3f summarize %.3f step %.3f ms %.2f Greads/s" % (r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], d["ms_per_step"], d["value"] / 1e9))
    except Exception as e:
        print(f, "failed", e)
P
}
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 $B > gpurun_out/b_main.json 2> gpurun_out/b.err
timeout 300 $B --mixed > gpurun_out/b_main_mixed.json 2>> gpurun_out/b.err
timeout 300 $B --validate > gpurun_out/b_main_validate.json 2>> gpurun_out/b.err
timeout 300 $B --mode views > gpurun_out/b_main_views.json 2>> gpurun_out/b.err
show gpurun_out/b_main.json gpurun_out/b_main_mixed.json gpurun_out/b_main_validate.json gpurun_out/b_main_views.json
ROUND_TAG=r02 bash scripts/gpu_stride_sweep.sh
V=blazeseq_b200/lib/variants
for tag in d6 ds2c5; do
  [ -f $V/lib_$tag.so ] || continue
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B > gpurun_out/b_$tag.json 2>> gpurun_out/b.err
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B --mixed > gpurun_out/b_${tag}_mixed.json 2>> gpurun_out/b.err
  show gpurun_out/b_$tag.json gpurun_out/b_${tag}_mixed.json
done
# one full capture of k_resolve (pack) per copy variant: launch 8 = a warm window of the second pass over the input
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 8 -c 1 -o gpurun_out/r02_prof_resolve_staged -f python bench.py --gib 4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_a.log 2>&1
BSQ_LIB=$V/lib_d6.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 8 -c 1 -o gpurun_out/r02_prof_resolve_direct -f python bench.py --gib 4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 8 -c 1 -o gpurun_out/r02_prof_resolve_staged_320 -f python bench.py --gib 4 --steps 1 --warmup 1 --no-cpu --no-e2e --id-digits 9 > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/b.err; ls -la gpurun_out | tail
