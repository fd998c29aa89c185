# N-rank benches through torchrun -- run with: gpurun --gpus N -- 'N=2 bash scripts/gpu_multi.sh'
#   default line (weak scaling, configs[1] per GPU, sub-results incl. the sharded configs[3]), the sharded mixed stream as
#   the headline (strong scaling), and the reference arm
mkdir -p gpurun_out
N=${N:-2}; R=${ROUND_TAG:-r02}
nvidia-smi -L > gpurun_out/multi_gpus.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $T --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/${R}_bench_n${N}.json 2> gpurun_out/bench_n${N}.err
cut -c1-3000 gpurun_out/${R}_bench_n${N}.json; tail -3 gpurun_out/bench_n${N}.err
timeout 900 $T --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --mixed --shard-stream > gpurun_out/${R}_bench_shard_n${N}.json 2>> gpurun_out/bench_n${N}.err
cut -c1-1500 gpurun_out/${R}_bench_shard_n${N}.json; tail -3 gpurun_out/bench_n${N}.err
timeout 600 $T --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${R}_bench_n${N}_reference.json 2>> gpurun_out/bench_n${N}.err
cut -c1-600 gpurun_out/${R}_bench_n${N}_reference.json
