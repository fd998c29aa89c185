# full parity, stream bench (parallel pread), sanitizer subset on the staged-copy kernel
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_gzip.py --gib 2.0 > gpurun_out/r01_bench_gzip.json 2> gpurun_out/bench_gzip.err; tail -c 1800 gpurun_out/r01_bench_gzip.json; tail -3 gpurun_out/bench_gzip.err
for tool in memcheck racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=800 -k "literal_streams_bit_exact or tiny_and_degenerate or all_newlines or id_strip or config1" > gpurun_out/sanitize_$tool.log 2>&1
echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$tool.log | tail -5
done
