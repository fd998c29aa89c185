# compute-sanitizer memcheck + racecheck + synccheck on small inputs: literal / degenerate streams, every record stride,
# buffer limits, the device inflater, FASTA (gpurun -- 'bash scripts/gpu_sanitize.sh')
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}; : > gpurun_out/${R}_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitize_small_$tool.log 2>&1
  echo "== $tool: scripts/sanitize_small.py rc=$?" >> gpurun_out/${R}_sanitizer.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok" gpurun_out/sanitize_small_$tool.log | tail -3 >> gpurun_out/${R}_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fasta.py -m gpu -x -q --timeout=500 -k "literal_streams_bit_exact or tiny_and_degenerate or all_newlines or id_strip or config1 or reference_literals or (record_stride_sweep and 9)" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: pytest subset rc=$?" >> gpurun_out/${R}_sanitizer.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3 >> gpurun_out/${R}_sanitizer.txt
done
cat gpurun_out/${R}_sanitizer.txt
