# round 2, visit r (2 GPUs): bench.py after the rank-agreement rework of the host-batch legs, N=1 and N=2
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-sub > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r_bench_n1.json')); print(d['value'], d['e2e'])"; tail -2 gpurun_out/r_bench_n1.err
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-sub > gpurun_out/r_bench_n2.json 2> gpurun_out/r_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r_bench_n2.json')); print(d['value'], d['e2e'])"; tail -2 gpurun_out/r_bench_n2.err
