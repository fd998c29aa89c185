# round 2, visit m: uniform inflate kernel with the funnel-shift reader (default), uniform v1, 1 / 8 warps per CTA; buffers kept across streams
set -x
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    for k in ("bgzf_device_inflate",):
        v = d[k]
        print(sys.argv[1], "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f launch %.3f wait_inflate %.3f pass %.3f reader %.3f" % (v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["launch_s"], v["wait_inflate_s"], v["gpu_pass_s"], v["reader_busy_s"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 120 python -m pytest tests -m gpu -q -x --timeout=100 -k "device_inflate_is_bit_exact" > gpurun_out/m_pytest0.log 2>&1 || { tail -15 gpurun_out/m_pytest0.log; echo "inflate test failed or hung: stopping"; exit 1; }
tail -2 gpurun_out/m_pytest0.log
timeout 400 python -m pytest tests -m gpu -q -x --timeout=200 -k "inflate or bgzf or stream_pipeline or writer or plain_gzip or whole_batches" > gpurun_out/m_pytest.log 2>&1; tail -4 gpurun_out/m_pytest.log
for r in 256 512; do
  timeout 200 python bench.py --gzip --gib 4 --region-mib $r > gpurun_out/m_gzip4_$r.json 2> gpurun_out/m.err; show gpurun_out/m_gzip4_$r.json
done
for v in uni1 w1 w8; do
  BSQ_LIB=blazeseq_b200/lib/variants/lib_$v.so timeout 200 python bench.py --gzip --gib 4 --region-mib 512 > gpurun_out/m_gzip4_512_$v.json 2>> gpurun_out/m.err; show gpurun_out/m_gzip4_512_$v.json
done
tail -3 gpurun_out/m.err
