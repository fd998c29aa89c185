# bench (value / e2e / roofline / cpu_baseline) + ncu launch list + one full capture per hot kernel
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -k "tiny_and_degenerate or shard_summaries" > gpurun_out/pytest_fix.log 2>&1; tail -5 gpurun_out/pytest_fix.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; tail -c 3000 gpurun_out/bench_${R}.json; tail -5 gpurun_out/bench_${R}.err
timeout 600 python bench.py --steps 5 --warmup 3 --mode views --no-cpu --no-e2e > gpurun_out/bench_${R}_views.json 2>> gpurun_out/bench_${R}.err; tail -c 1500 gpurun_out/bench_${R}_views.json
timeout 600 python bench.py --steps 5 --warmup 3 --validate --no-cpu --no-e2e > gpurun_out/bench_${R}_validate.json 2>> gpurun_out/bench_${R}.err; tail -c 1500 gpurun_out/bench_${R}_validate.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}.err; tail -c 1200 gpurun_out/bench_${R}_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch_${R}.log 2>&1
tail -3 gpurun_out/ncu_launch_${R}.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 6 -c 2 -o gpurun_out/prof_resolve_${R} -f python bench.py --steps 1 --warmup 1 --gib 4 --no-cpu --no-e2e > gpurun_out/ncu_resolve_${R}.log 2>&1
tail -3 gpurun_out/ncu_resolve_${R}.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_summarize -s 6 -c 2 -o gpurun_out/prof_summarize_${R} -f python bench.py --steps 1 --warmup 1 --gib 4 --no-cpu --no-e2e > gpurun_out/ncu_summarize_${R}.log 2>&1
tail -3 gpurun_out/ncu_summarize_${R}.log
ls -la gpurun_out
