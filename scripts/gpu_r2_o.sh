# round 2, visit o: compute-sanitizer over the new kernels (uniform inflate, device writer), inflate prefetch A/B
set -x
mkdir -p gpurun_out
R=r02b; : > gpurun_out/${R}_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitize_small_$tool.log 2>&1
  echo "== $tool: scripts/sanitize_small.py rc=$?" >> gpurun_out/${R}_sanitizer.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok" gpurun_out/sanitize_small_$tool.log | tail -3 >> gpurun_out/${R}_sanitizer.txt
done
timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=350 -k "device_inflate_is_bit_exact or device_writer" > gpurun_out/sanitize_new_memcheck.log 2>&1
echo "== memcheck: pytest device_inflate_is_bit_exact + device_writer rc=$?" >> gpurun_out/${R}_sanitizer.txt
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_new_memcheck.log | tail -3 >> gpurun_out/${R}_sanitizer.txt
cat gpurun_out/${R}_sanitizer.txt
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    for k in ("bgzf_device_inflate",):
        v = d[k]
        print(sys.argv[1], "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f wait_inflate %.3f pass %.3f regions %d" % (v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["wait_inflate_s"], v["gpu_pass_s"], v["regions"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 200 python bench.py --gzip --gib 4 > gpurun_out/o_gzip4_512.json 2> gpurun_out/o.err; show gpurun_out/o_gzip4_512.json
BSQ_LIB=blazeseq_b200/lib/variants/lib_nopf.so timeout 200 python bench.py --gzip --gib 4 > gpurun_out/o_gzip4_512_nopf.json 2>> gpurun_out/o.err; show gpurun_out/o_gzip4_512_nopf.json
timeout 200 python bench.py --gzip --gib 4 > gpurun_out/o_gzip4_512_b.json 2>> gpurun_out/o.err; show gpurun_out/o_gzip4_512_b.json
tail -3 gpurun_out/o.err
