mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    for k in ("bgzf_device_inflate",):
        v = d[k]
        print(k, "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f pass %.3f reader %.3f" % (v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["gpu_pass_s"], v["reader_busy_s"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 90 python -m pytest tests -m gpu -q -x --timeout=60 -k "device_inflate_is_bit_exact" > gpurun_out/pytest_gpu0.log 2>&1 || { tail -5 gpurun_out/pytest_gpu0.log; echo "inflate test failed or hung: stopping"; exit 1; }
tail -2 gpurun_out/pytest_gpu0.log
timeout 200 python -m pytest tests -m gpu -q -x --timeout=100 -k "inflate or bgzf or stream_pipeline or writer" > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for r in 256 512; do
  timeout 150 python bench.py --gzip --gib 4 --region-mib $r > gpurun_out/j_gzip4_$r.json 2> gpurun_out/j.err; show gpurun_out/j_gzip4_$r.json
done
tail -3 gpurun_out/j.err
