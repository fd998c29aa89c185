# round 2, visit q (2 GPUs): the driver's N=2 line after this round's changes (NUMA binding, pipelined host batches), racecheck of
# the inflate kernel with the warp barrier in front of the table rebuild
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/q_bench_n2.json 2> gpurun_out/q_bench_n2.err
cut -c1-2500 gpurun_out/q_bench_n2.json; tail -3 gpurun_out/q_bench_n2.err
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/q_sanitize_small_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|sanitize_small ok" gpurun_out/q_sanitize_small_racecheck.log
