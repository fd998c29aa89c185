# the committed evidence: bench line, launch list of the same command, one full capture per hot kernel
set -x
mkdir -p gpurun_out
R=${ROUND_TAG:-r01}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${R}_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --mode views --no-cpu --no-e2e > gpurun_out/${R}_bench_views.json 2>> gpurun_out/${R}_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --validate --no-cpu --no-e2e > gpurun_out/${R}_bench_validate.json 2>> gpurun_out/${R}_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --mixed --no-cpu --no-e2e > gpurun_out/${R}_bench_mixed.json 2>> gpurun_out/${R}_bench.err
BSQ_SINGLE_PASS=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_single_pass.json 2>> gpurun_out/${R}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${R}_ncu_launch.log 2>&1
# launch 13 of k_resolve / k_summarize = window 0 of the first timed step (steps: 3 warm-up + ...; 6 windows per step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 18 -c 1 -o gpurun_out/${R}_prof_resolve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${R}_ncu_resolve.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_summarize -s 18 -c 1 -o gpurun_out/${R}_prof_summarize -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${R}_ncu_summarize.log 2>&1
tail -2 gpurun_out/${R}_bench.err
ls -la gpurun_out | tail -15
