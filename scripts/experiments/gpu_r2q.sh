# list barrier first, tile bytes waited for where they are first read: parity + benches
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 $B > gpurun_out/q_batches.json 2>> gpurun_out/ab.err; show gpurun_out/q_batches.json "batches"
timeout 600 $B --mode views > gpurun_out/q_views.json 2>> gpurun_out/ab.err; show gpurun_out/q_views.json "views"
timeout 600 $B --validate > gpurun_out/q_validate.json 2>> gpurun_out/ab.err; show gpurun_out/q_validate.json "validate"
timeout 600 $B --mixed > gpurun_out/q_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/q_mixed.json "mixed"
tail -3 gpurun_out/ab.err
