# round 2, final visit: smoke, the whole GPU suite, the driver's bench lines (own arm, reference arm), bench --gzip, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/p_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/p_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; tail -2 gpurun_out/p_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/p_pytest_gpu.log 2>&1; tail -5 gpurun_out/p_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_reference.json 2> gpurun_out/p_ref.err; tail -c 600 gpurun_out/p_bench_reference.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; tail -c 2500 gpurun_out/p_bench.json; tail -3 gpurun_out/p_bench.err
timeout 400 python bench.py --gzip --gib 4 > gpurun_out/p_bench_gzip.json 2> gpurun_out/p_gzip.err; tail -c 1200 gpurun_out/p_bench_gzip.json; tail -3 gpurun_out/p_gzip.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > gpurun_out/p_ncu_launch.log 2>&1
tail -3 gpurun_out/p_launches.csv | cut -c1-200
