# round 2, last visit: the whole GPU suite and the driver's lines on the final tree
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; tail -1 gpurun_out/s_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/s_pytest_gpu.log 2>&1; tail -3 gpurun_out/s_pytest_gpu.log
timeout 400 python bench.py --gzip --gib 4 > gpurun_out/s_bench_gzip.json 2> gpurun_out/s_gzip.err; tail -3 gpurun_out/s_gzip.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -3 gpurun_out/s_bench.err
python - <<'P'
import json
g = json.load(open("gpurun_out/s_bench_gzip.json"))["gzip"]
for k in ("bgzf_device_inflate", "gzip_parallel_host_threads", "bgzf_host_threads", "plain_file", "cpu_zlib_all_threads_inflate_only"):
    print(k, round(g[k]["uncompressed_gb_per_s"], 2), round(g[k]["wall_s"], 3))
d = json.load(open("gpurun_out/s_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["host_batch"]["ms_per_step"], d["e2e"]["host_batch_pipelined"]["ms_per_step"], d["sub_results"]["configs[4]"])
P
