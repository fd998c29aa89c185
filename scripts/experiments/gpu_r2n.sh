# validation screen in k_summarize (clean tiles take the list in the validating k_resolve kernels): parity + A/B
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 $B --validate > gpurun_out/r01_bench_validate.json 2>> gpurun_out/ab.err; show gpurun_out/r01_bench_validate.json "validate screen"
BSQ_NO_LIST_HANDOFF=1 timeout 600 $B --validate > gpurun_out/ab_nolist.json 2>> gpurun_out/ab.err; show gpurun_out/ab_nolist.json "validate no_handoff"
timeout 600 $B > gpurun_out/ab_list.json 2>> gpurun_out/ab.err; show gpurun_out/ab_list.json "batches"
timeout 600 $B --mode views --validate > gpurun_out/ab_vv.json 2>> gpurun_out/ab.err; show gpurun_out/ab_vv.json "views validate"
tail -3 gpurun_out/ab.err
