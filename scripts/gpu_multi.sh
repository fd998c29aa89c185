# N-rank bench through torchrun (weak scaling) -- run with: gpurun --gpus N -- 'N=2 bash scripts/gpu_multi.sh'
set -x
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r01_bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 2500 gpurun_out/r01_bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r01_bench_n${N}_reference.json 2>> gpurun_out/bench_n${N}.err
tail -c 600 gpurun_out/r01_bench_n${N}_reference.json
