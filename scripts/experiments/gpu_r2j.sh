# 2-rank torchrun bench (weak scaling) + reference arm under torchrun + the consumer test
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -k "quality_sums or shard" > gpurun_out/pytest_new.log 2>&1; tail -4 gpurun_out/pytest_new.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01_bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 1800 gpurun_out/r01_bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r01_bench_n${N}_reference.json 2>> gpurun_out/bench_n${N}.err
tail -c 400 gpurun_out/r01_bench_n${N}_reference.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --mixed --no-e2e --no-cpu > gpurun_out/r01_bench_n${N}_mixed.json 2>> gpurun_out/bench_n${N}.err
tail -c 600 gpurun_out/r01_bench_n${N}_mixed.json
