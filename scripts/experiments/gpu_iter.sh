# quick iteration: parity suite (fail fast) + short benches
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q -x --timeout=900 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; python -c "
import json;d=json.load(open('gpurun_out/bench_iter.json'));r=d['roofline']
print('batches value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f  e2e %s'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac'],d['e2e'] and '%.3g'%d['e2e']['value']))"
tail -3 gpurun_out/bench_iter.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mode views > gpurun_out/bench_iter_views.json 2>> gpurun_out/bench_iter.err; python -c "
import json;d=json.load(open('gpurun_out/bench_iter_views.json'));r=d['roofline']
print('views   value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --validate > gpurun_out/bench_iter_val.json 2>> gpurun_out/bench_iter.err; python -c "
import json;d=json.load(open('gpurun_out/bench_iter_val.json'));r=d['roofline']
print('validate value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"
