# round 2: what the start rotation / lane interleave of the staged copy cost at a 319-byte stride; full captures
mkdir -p gpurun_out
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
timeout 300 $B > gpurun_out/c_main.json 2> gpurun_out/c.err
timeout 300 $B --id-digits 9 > gpurun_out/c_main_320.json 2>> gpurun_out/c.err
show gpurun_out/c_main.json gpurun_out/c_main_320.json
for tag in r0 i0 r0i0; do
  [ -f $V/lib_$tag.so ] || continue
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B > gpurun_out/c_$tag.json 2>> gpurun_out/c.err
  BSQ_LIB=$V/lib_$tag.so timeout 300 $B --id-digits 9 > gpurun_out/c_${tag}_320.json 2>> gpurun_out/c.err
  show gpurun_out/c_$tag.json gpurun_out/c_${tag}_320.json
done
# full captures: one window of 1.9 GiB; launches per pass: 1 k_summarize + 1 k_resolve; 3 warm-up passes
N="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
P="python bench.py --gib 1.9 --steps 1 --warmup 3 --no-cpu --no-e2e"
timeout 600 $N -k regex:k_resolve -o gpurun_out/r02_prof_resolve_staged $P > gpurun_out/ncu_a.log 2>&1
timeout 600 $N -k regex:k_resolve -o gpurun_out/r02_prof_resolve_staged_320 $P --id-digits 9 > gpurun_out/ncu_c.log 2>&1
timeout 600 $N -k regex:k_summarize -o gpurun_out/r02_prof_summarize $P > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/c.err; ls -la gpurun_out/*.ncu-rep
