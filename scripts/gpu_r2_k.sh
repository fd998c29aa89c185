# round 2, visit k: the parallel gzip decoder in the stream pipeline, BSQ_WANT_WHOLE_BATCHES + HostBatchPipeline, full suite, bench lines
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/k_gpu.txt; nproc >> gpurun_out/k_gpu.txt
timeout 300 python -m pytest tests -m gpu -q -x --timeout=200 -k "plain_gzip or whole_batches or stream_pipeline or file_and_gzip" > gpurun_out/k_pytest_new.log 2>&1; tail -5 gpurun_out/k_pytest_new.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; tail -2 gpurun_out/k_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/k_pytest_gpu.log 2>&1; tail -8 gpurun_out/k_pytest_gpu.log
timeout 400 python bench.py --gzip --gib 4 > gpurun_out/k_bench_gzip.json 2> gpurun_out/k_gzip.err; tail -c 1500 gpurun_out/k_bench_gzip.json; tail -3 gpurun_out/k_gzip.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; tail -c 3000 gpurun_out/k_bench.json; tail -3 gpurun_out/k_bench.err
