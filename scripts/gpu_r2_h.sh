mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    v = d["bgzf_device_inflate"]
    print(sys.argv[1].split("/")[-1], "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f pass %.3f" % (v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["gpu_pass_s"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 600 python -m pytest tests -m gpu -q -x --timeout=900 -k "inflate or bgzf" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for r in 256 1024; do
  timeout 600 python bench.py --gzip --gib 2 --region-mib $r > gpurun_out/h_main_$r.json 2> gpurun_out/h.err; show gpurun_out/h_main_$r.json
  BSQ_LIB=blazeseq_b200/lib/variants/lib_g8.so timeout 600 python bench.py --gzip --gib 2 --region-mib $r > gpurun_out/h_g8_$r.json 2>> gpurun_out/h.err; show gpurun_out/h_g8_$r.json
done
BSQ_LIB=blazeseq_b200/lib/variants/lib_g8.so timeout 600 python -m pytest tests -m gpu -q -x --timeout=900 -k "inflate or bgzf" > gpurun_out/pytest_g8.log 2>&1; tail -3 gpurun_out/pytest_g8.log
tail -3 gpurun_out/h.err
