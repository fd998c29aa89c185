# one full ncu capture per hot kernel on a 2 GiB input (one window), batches mode
set -x
mkdir -p gpurun_out
T=${TAG:-cur}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 3 -c 1 -o gpurun_out/prof_resolve_$T -f python bench.py --steps 1 --warmup 1 --gib 1.9 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_resolve_$T.log 2>&1
tail -2 gpurun_out/ncu_resolve_$T.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_summarize -s 3 -c 1 -o gpurun_out/prof_summarize_$T -f python bench.py --steps 1 --warmup 1 --gib 1.9 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_summarize_$T.log 2>&1
tail -2 gpurun_out/ncu_summarize_$T.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mode views > gpurun_out/bench_iter_views.json 2>> gpurun_out/bench_iter.err; python -c "
import json;d=json.load(open('gpurun_out/bench_iter_views.json'));r=d['roofline']
print('views   value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"
