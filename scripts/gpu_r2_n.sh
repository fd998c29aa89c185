# round 2, visit n: device FASTQ writer, ramped first regions, 8 warps per CTA in the inflate kernel, ncu of a full-size inflate launch
set -x
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    for k in ("bgzf_device_inflate",):
        v = d[k]
        print(sys.argv[1], "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f launch %.3f wait_inflate %.3f pass %.3f reader %.3f wait_reader %.3f regions %d" % (v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["launch_s"], v["wait_inflate_s"], v["gpu_pass_s"], v["reader_busy_s"], v["caller_wait_reader_s"], v["regions"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 120 python -m pytest tests -m gpu -q -x --timeout=100 -k "device_inflate_is_bit_exact or device_writer" > gpurun_out/n_pytest0.log 2>&1 || { tail -30 gpurun_out/n_pytest0.log; echo "first tests failed or hung"; }
tail -2 gpurun_out/n_pytest0.log
timeout 400 python -m pytest tests -m gpu -q --timeout=200 -k "inflate or bgzf or stream_pipeline or writer or plain_gzip or whole_batches" > gpurun_out/n_pytest.log 2>&1; tail -6 gpurun_out/n_pytest.log
for r in 256 512; do
  timeout 200 python bench.py --gzip --gib 4 --region-mib $r > gpurun_out/n_gzip4_$r.json 2> gpurun_out/n.err; show gpurun_out/n_gzip4_$r.json
done
tail -3 gpurun_out/n.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate_members -s 3 -c 1 -f -o gpurun_out/r02_prof_k_inflate python bench.py --gzip --gib 1 --region-mib 256 > gpurun_out/ncu_k_inflate.log 2>&1
timeout 120 python scripts/profile_summary.py gpurun_out/r02_prof_k_inflate.ncu-rep k_inflate > gpurun_out/r02_k_inflate_members.txt 2>&1
head -40 gpurun_out/r02_k_inflate_members.txt | cut -c1-160
ls -la gpurun_out/*.ncu-rep
