# compute-sanitizer memcheck + racecheck + synccheck on small inputs (smoke + a few parity tests)
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
timeout 1200 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=1000 -k "literal_streams_bit_exact or tiny_and_degenerate or all_newlines or id_strip or config1" > gpurun_out/sanitize_$tool.log 2>&1
echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$tool.log | tail -5
done
