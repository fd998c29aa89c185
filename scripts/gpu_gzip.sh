# configs[4] on the GPU box: the gzip / BGZF parity tests, `bench.py --gzip` at the given region sizes (and, with VARIANTS="a b",
# through blazeseq_b200/lib/variants/lib_<tag>.so builds of scripts/build_variants.sh), one ncu capture of a full-size inflate launch
#   gpurun -- 'REGIONS="256 512" VARIANTS="w4" bash scripts/gpu_gzip.sh'
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))["gzip"]
    for k in ("bgzf_device_inflate", "gzip_parallel_host_threads"):
        v = d[k]
        print(sys.argv[1], k, "region", d["region_mib"], "MiB: %.2f GB/s wall %.3f s h2d %.3f inflate %.3f wait_inflate %.3f reader %.3f regions %d" % (
            v["uncompressed_gb_per_s"], v["wall_s"], v["h2d_compressed_s"], v["inflate_kernels_s"], v["wait_inflate_s"], v["reader_busy_s"], v["regions"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
timeout 400 python -m pytest tests -m gpu -q -x --timeout=200 -k "inflate or bgzf or stream_pipeline or writer or plain_gzip or whole_batches" > gpurun_out/gzip_pytest.log 2>&1; tail -3 gpurun_out/gzip_pytest.log
for r in ${REGIONS:-512}; do
  timeout 300 python bench.py --gzip --gib ${GIB:-4} --region-mib $r > gpurun_out/${R}_bench_gzip_$r.json 2> gpurun_out/gzip.err; show gpurun_out/${R}_bench_gzip_$r.json
  for v in ${VARIANTS:-}; do
    BSQ_LIB=blazeseq_b200/lib/variants/lib_$v.so timeout 300 python bench.py --gzip --gib ${GIB:-4} --region-mib $r > gpurun_out/${R}_bench_gzip_${r}_$v.json 2>> gpurun_out/gzip.err; show gpurun_out/${R}_bench_gzip_${r}_$v.json
  done
done
tail -3 gpurun_out/gzip.err
# the 4th inflate launch of a stream with 256 MiB regions is the first full-size one (the first three regions are 1/8, 1/4, 1/2)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate_members -s 3 -c 1 -f -o gpurun_out/${R}_prof_k_inflate python bench.py --gzip --gib 1 --region-mib 256 > gpurun_out/ncu_k_inflate.log 2>&1
timeout 120 python scripts/profile_summary.py gpurun_out/${R}_prof_k_inflate.ncu-rep k_inflate > gpurun_out/${R}_k_inflate_members.txt 2>&1
head -30 gpurun_out/${R}_k_inflate_members.txt | cut -c1-150
