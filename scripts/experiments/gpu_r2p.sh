# sparse views passes (a listed tile is not loaded at all): parity + A/B
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 $B --mode views > gpurun_out/r01_bench_views.json 2>> gpurun_out/ab.err; show gpurun_out/r01_bench_views.json "views sparse"
timeout 600 $B > gpurun_out/ab_list.json 2>> gpurun_out/ab.err; show gpurun_out/ab_list.json "batches"
tail -3 gpurun_out/ab.err
