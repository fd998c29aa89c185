# BGZF block-parallel inflate: parity + configs[4] bench; 1-GPU quick bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "stream or bgzf or file_and_gzip or python_binding or native" > gpurun_out/pytest_stream.log 2>&1; tail -15 gpurun_out/pytest_stream.log
timeout 900 python scripts/bench_gzip.py --gib 2.0 > gpurun_out/r01_bench_gzip.json 2> gpurun_out/bench_gzip.err; tail -c 2500 gpurun_out/r01_bench_gzip.json; tail -3 gpurun_out/bench_gzip.err
