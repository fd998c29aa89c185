# newline-list hand-off (k_summarize -> k_resolve): parity + A/B against BSQ_NO_LIST_HANDOFF=1 on the same box
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for m in "" "--mode views" "--mixed" "--validate"; do
timeout 600 $B $m > gpurun_out/ab_list.json 2>> gpurun_out/ab.err; show gpurun_out/ab_list.json "handoff $m"
BSQ_NO_LIST_HANDOFF=1 timeout 600 $B $m > gpurun_out/ab_nolist.json 2>> gpurun_out/ab.err; show gpurun_out/ab_nolist.json "no_handoff $m"
done
tail -3 gpurun_out/ab.err
