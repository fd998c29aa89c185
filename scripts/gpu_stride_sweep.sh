# record-stride sweep of k_resolve (stride = 2 L + id digits + 11): one 2 GiB window per point
mkdir -p gpurun_out
OUT=gpurun_out/${ROUND_TAG:-r02}_stride_sweep.jsonl; : > $OUT
for spec in "147 8" "148 8" "149 8" "150 8" "150 9" "151 8" "151 9" "152 8" "153 8" "154 8" "155 8" "134 8" "166 8"; do
  set -- $spec
  timeout 300 python bench.py --gib 2 --steps 5 --warmup 3 --no-cpu --no-e2e --read-len $1 --id-digits $2 ${BENCH_ARGS:-} >> $OUT 2>> gpurun_out/stride.err
done
python - <<'P'
import json, os
for line in open(os.environ.get("OUT", "gpurun_out/%s_stride_sweep.jsonl" % os.environ.get("ROUND_TAG", "r02"))):
    d = json.loads(line); r = d["roofline"]
    print("stride %.0f  k_resolve %.4f ms  frac %.3f  summarize %.3f ms  step %.3f ms  %.2f Greads/s" % (
        d["config"]["record_bytes"], r["avg_launch_ms"], r["frac"], r["summarize_ms_per_step"], d["ms_per_step"], d["value"] / 1e9))
P
