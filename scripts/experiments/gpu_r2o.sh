# verification of HEAD: full parity (incl. the screen test), two-pass fallbacks, sanitizer subset
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
BSQ_NO_LIST_HANDOFF=1 timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "corpus or random or degenerate or all_newlines or long_reads or screen or config1" > gpurun_out/pytest_nolist.log 2>&1; tail -2 gpurun_out/pytest_nolist.log
for tool in memcheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=800 -k "literal_streams_bit_exact or tiny_and_degenerate or all_newlines or id_strip or config1" > gpurun_out/sanitize_$tool.log 2>&1
echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$tool.log | tail -4
done
