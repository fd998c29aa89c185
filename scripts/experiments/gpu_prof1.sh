set -x
mkdir -p gpurun_out
T=${TAG:-cur}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN:-k_resolve} -s 3 -c 1 -o gpurun_out/prof_${KERN:-k_resolve}_$T -f python bench.py --steps 1 --warmup 1 --gib 1.9 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_$T.log 2>&1
tail -2 gpurun_out/ncu_$T.log
