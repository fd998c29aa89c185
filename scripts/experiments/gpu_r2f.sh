# parity + quick benches of the round's default build
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  tail %.2f step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['tail_rebase_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 $B > gpurun_out/ab_def.json 2> gpurun_out/ab.err; show gpurun_out/ab_def.json batches
timeout 600 $B --mode views > gpurun_out/ab_def_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_views.json views
timeout 600 $B --validate > gpurun_out/ab_def_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_val.json validate
timeout 600 $B --mixed > gpurun_out/ab_def_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_mixed.json mixed
tail -5 gpurun_out/ab.err
